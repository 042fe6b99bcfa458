timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/c9_pytest.log
tail -12 gpurun_out/c9_pytest.log
drivers/scripts -n 256 -r 10 2>&1 | tee gpurun_out/c9_scripts.log
