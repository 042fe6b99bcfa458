// bk_stencil_remote.cu -- the marching kernels with THE EXCHANGE INSIDE THE SWEEP (bk_stencil_advance_remote).
//
// The reference exchanges ghost zones as a step of its own: pack / MPI / unpack on the host (weak/main.cu:251-272), here
// so far one pull kernel that copies the neighbours' skin bricks into the own ghost bricks (bk_exchange.cu), after
// which the first sweep of the period reads them back.  With peer mappings the copy is not needed at all: the sweep's
// producer warps can bulk-copy a ghost brick's planes into shared memory straight from the NEIGHBOUR's storage -- the
// transfer happens brick by brick inside the kernel that consumes it, overlapped with the math of every other tile, and
// the 104 MB ghost write and re-read per period disappear together with the pull kernel that time-shares SMs with the
// sweep.  Only the launches that read freshly exchanged ghosts take this path (the REST half of a period's first pass).
//
// This translation unit IS bk_stencil_tiled.cu, compiled with BK_REMOTE_TU: the marching body gains one statement in its
// producer (ghost id -> address through the RemoteArgs table) and its own kernels / entry point; the kernels of the
// other translation unit are not touched by its existence.
#define BK_REMOTE_TU 1
#include "bk_stencil_tiled.cu"
