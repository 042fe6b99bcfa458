python tools/sweep_bench.py --steps 2 --stencils mpi7pt --variants 0,11,12,13,14,15,16 --reps 20 2>&1 | tee gpurun_out/c10_abl.log
