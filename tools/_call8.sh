nvidia-smi topo -m | head -12 > gpurun_out/c15_topo.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4 > gpurun_out/c15_pytest.log
tail -3 gpurun_out/c15_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c15_bench8.json 2> gpurun_out/c15_bench8.err
tail -2 gpurun_out/c15_bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/c15_ref8.json 2> gpurun_out/c15_ref8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py 2>/dev/null | tail -1 | tee gpurun_out/c15_pcie8.log
python -c "
import json
d=json.loads(open('gpurun_out/c15_bench8.json').read().strip().splitlines()[-1])
print('7pt',d['value'], d['ms_per_step'], d['parity']['ok'], d['parity']['max_rel'])
for k,v in d['others'].items():
    if 'GStencil/s' in v: print(k, round(v['GStencil/s'],1), round(v['ms_per_step'],4), v['parity']['ok'])
    else: print(k, json.dumps(v)[:900])
print(d['e2e'])
r=json.loads(open('gpurun_out/c15_ref8.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
"
