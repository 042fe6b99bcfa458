timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c5_pytest.log
timeout 600 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
tail -3 gpurun_out/c5_pytest.log; tail -5 gpurun_out/c5_bench.err; wc -c gpurun_out/c5_bench.json
