timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c6_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c6_bench2.json 2> gpurun_out/c6_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c6_ref2.json 2> gpurun_out/c6_ref2.err
tail -3 gpurun_out/c6_pytest.log; tail -3 gpurun_out/c6_bench2.err; wc -c gpurun_out/c6_bench2.json gpurun_out/c6_ref2.json
