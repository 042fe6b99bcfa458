timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c14_bench2.json 2> gpurun_out/c14_bench2.err
tail -3 gpurun_out/c14_bench2.err
python -c "
import json
d=json.loads(open('gpurun_out/c14_bench2.json').read().strip().splitlines()[-1])
print(d['value']); print(json.dumps(d['others']['strong'],indent=0)[:1500]); print(json.dumps(d['others'].get('array_layout_baseline'))[:600])
"
