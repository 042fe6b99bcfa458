python tools/sweep_bench.py --steps 1 --stencils mpi25pt --variants 0,2,3,4 --reps 20 2>&1 | tee gpurun_out/c4_25.log
python tools/sweep_bench.py --steps 1 --stencils mpi25pt --variants 0,2,3,4 --reps 20 --full 2>&1 | tee gpurun_out/c4_25full.log
