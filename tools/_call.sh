timeout 900 python -m pytest tests/test_gpu_array.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c7_pytest.log
tail -15 gpurun_out/c7_pytest.log
drivers/weak -s 512,512,512 -I 10 -g 1 -S mpi7pt 2>&1 | tee gpurun_out/c7_weak7.log | grep -E "^Arr|^Bri|perf|Arr =="
drivers/weak -s 512,512,512 -I 10 -g 1 -S mpi25pt 2>&1 | tee gpurun_out/c7_weak25.log | grep -E "^Arr|^Bri|perf|Arr =="
drivers/weak -s 512,512,512 -I 10 -g 1 -S mpi125pt 2>&1 | tee gpurun_out/c7_weak125.log | grep -E "^Arr|^Bri|perf|Arr =="
