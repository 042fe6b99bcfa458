#!/bin/bash
# What the round-2 session could not run (it ended without GPU time): the first contact of everything written after the
# budget was spent, in the order that loses least if something goes wrong.  One GPU; every step under its own timeout;
# results in gpurun_out/.   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/fc_$name.log" 2>&1; echo "rc=$? ($name)"; tail -3 "gpurun_out/fc_$name.log"; }
run smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
run trial 300 python tools/composed_trial.py
run handshake2 200 python tests/handshake_case.py --ranks 2
run handshake4 200 python tests/handshake_case.py --ranks 4
run direct 400 python tools/direct_exchange_trial.py
run pytest 1200 python -m pytest tests -m gpu -x -q
run bench1 600 python bench.py --steps 20 --warmup 5
for v in staged composed wide; do
  run sweep_$v 120 env BK_FUSED_VARIANT=$v python tools/sweep_bench.py --steps 2 --stencils mpi7pt --reps 20
done
# one full capture of the composed kernel (never under a profiler for a number: this is for the counters)
run ncu_composed 600 env BK_FUSED_VARIANT=composed ncu --set full --clock-control none --import-source on -k regex:k_star_capped -s 3 -c 1 \
    -o gpurun_out/r03_composed python tools/sweep_bench.py --steps 2 --stencils mpi7pt --reps 1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/fc_bench1.log").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "fused", d["fused_kernel"].get("selected"), d["fused_kernel"].get("launch_ms"))
    print("loop", d["loop_options"]); print("roofline", d["roofline"]["frac"], d["roofline"]["hbm_frac"], "parity", d["parity"]["ok"])
except Exception as exc:
    print("bench line not readable:", exc)
PY
