#!/bin/bash
# Round 2, call 3 (1 GPU): whole GPU suite on the new defaults (v3 kernel, GEMM TRSM for tall panels, new rowscale),
# the headline bench (C4 at N = 1 with the secondary C2), launch list of a C4h fit.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02c3_tests.log
echo "== 2. bench C4 (N = 1)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c3_bench_c4_1gpu.json 2> gpurun_out/r02c3_bench_c4_1gpu.err
tail -c 300 gpurun_out/r02c3_bench_c4_1gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c3_bench_c4_1gpu.json").read().strip().splitlines()[-1])
    print("C4 value", d["value"], "e2e", d["e2e"]["value"], "third_n3", d.get("third_n3_tflops"), "phases", d["phases_ms"])
    r = d["roofline"]; print("roofline achieved", r["achieved"], "peak", r["peak"], "frac", r["frac"], "kernel_ms", r["kernel_ms_per_step"], "nominal frac", r.get("frac_of_nominal_4500"))
    print("parity", d["parity"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
    c2 = d.get("c2"); print("c2", c2 and {k: c2[k] for k in ("value", "e2e", "third_n3_tflops", "parity")})
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 3. reference arm (bounded)"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02c3_bench_ref.json 2> gpurun_out/r02c3_bench_ref.err
cat gpurun_out/r02c3_bench_ref.json | cut -c1-400; tail -4 gpurun_out/r02c3_bench_ref.err
echo "== 4. launch list C4h"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02c3_launches_c4h.csv python tools/fit_once.py 32768 64 1 > /dev/null 2>&1
for la in 1 2; do AGP_LOOKAHEAD=$la timeout 120 python tools/fit_once.py 4096 8 8; done
timeout 120 python tools/fit_once.py 32768 64 3
ls -la gpurun_out | grep r02c3
