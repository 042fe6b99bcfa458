timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/c11_pytest.log
tail -6 gpurun_out/c11_pytest.log
python tools/sweep_bench.py --steps 2 --stencils mpi7pt --variants 0,13 --reps 20 2>&1 | tee gpurun_out/c11_fused.log
drivers/strong -d 1024 -s 64 -I 10 -g 1 -S mpi7pt 2>&1 | grep -E "perf|calc|call|wait" | tee gpurun_out/c11_strong.log
drivers/strong -d 1024 -s 64 -I 10 -g 1 -S mpi7pt -O 2>&1 | grep -E "perf|calc|call|wait" | tee -a gpurun_out/c11_strong.log
