#!/bin/bash
# Round 2, call 4 (1 GPU): the generalised tcgen05 product (fp32 / fp64 operands, rectangular, accumulate), the fp32
# factorisation and the tensor-core forward substitution; C3 at full size.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. new kernel tests"
timeout 600 python -m pytest tests/test_gpu_ozaki.py -q -m gpu -k "general_product or fp32_syrk" 2>&1 | tail -15 | tee gpurun_out/r02c4_t_kernel.log
echo "== 2. in-situ fp32 / predict tests"
timeout 600 python -m pytest tests/test_gpu_tcgen05_insitu.py -q -m gpu -k "fp32 or predict" 2>&1 | tail -15 | tee gpurun_out/r02c4_t_insitu.log
echo "== 3. whole GPU suite"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02c4_t_all.log
echo "== 4. bench C3 (tensor path, then FFMA path)"
timeout 600 python bench.py --workload C3 --steps 5 --warmup 3 > gpurun_out/r02c4_bench_c3.json 2> gpurun_out/r02c4_bench_c3.err
tail -c 400 gpurun_out/r02c4_bench_c3.err
AGP_FP32_MODE=0 timeout 600 python bench.py --workload C3 --steps 3 --warmup 3 > gpurun_out/r02c4_bench_c3_ffma.json 2> gpurun_out/r02c4_bench_c3_ffma.err
python - <<'PY'
import json
for t in ("c3", "c3_ffma"):
    try:
        d = json.loads(open("gpurun_out/r02c4_bench_%s.json" % t).read().strip().splitlines()[-1])
        print(t, "value", d["value"], "e2e", d["e2e"]["value"], "phases", d["phases_ms"], "parity", d["parity"], "roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "kernel_ms_per_step")})
    except Exception as e:
        print(t, "parse failed", e)
PY
