#!/bin/bash
# Round 2, call 8 (1 GPU): regression of the day's changes + C5-shaped VFE at shard size + the 1-rank distributed probe
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r02c8_tests.log
echo "== 2. 1-rank distributed path, host vs device pointers"
timeout 300 python tools/dist1_probe.py 32768 64 3 2>&1 | tail -14 | tee gpurun_out/r02c8_dist1.log
echo "== 3. C5-shaped VFE (fp32, D=16, M=8192) at N = 62500 (half a rank's shard at 8 GPUs): tensor vs FFMA"
timeout 600 python bench.py --workload C5 --n 62500 --steps 2 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-500
AGP_FP32_MODE=0 timeout 900 python bench.py --workload C5 --n 62500 --steps 1 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-500
echo "== 4. C5 parity-size (N=20000, M=2500) full line"
timeout 600 python bench.py --workload C5 --n 20000 --steps 3 --warmup 3 > gpurun_out/r02c8_bench_c5_n20000.json 2> gpurun_out/r02c8_bench_c5.err; tail -c 300 gpurun_out/r02c8_bench_c5.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c8_bench_c5_n20000.json").read().strip().splitlines()[-1])
    print("C5@20000 value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"], "roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac")})
except Exception as e:
    print("parse failed", e)
PY
