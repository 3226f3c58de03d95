#!/bin/bash
# Round 2, call 2: the restructured trailing-update kernel (v3: uniform-datapath issue loops, interleaved slice layout,
# exact int64 + magic-number drain).  Parity first, then the probe of its phases / variants and one ncu pass.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. parity: kernel tests, in-situ fits, variants"
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_tcgen05_insitu.py tests/test_gpu_variants_grad_vfecov.py -q -m gpu -k "ozaki or insitu or tcgen05 or strip or variant" 2>&1 | tail -12 | tee gpurun_out/r02c2_tests.log
echo "== 2. probe (v3 default, variants, v2 for the record) + C4h fits"
PROBE_CLUSTER=1 PROBE_FIT=1 timeout 400 python tools/ozaki_probe.py > gpurun_out/r02c2_probe_stdout.json 2> gpurun_out/r02c2_probe.err
cp gpurun_out/ozaki_probe.json gpurun_out/r02c2_ozaki_probe.json 2>/dev/null
tail -c 600 gpurun_out/r02c2_probe.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c2_ozaki_probe.json"))
    a = d["partA"]
    print("fixed_ms", a["fixed_ms"])
    for k, v in a["modes"].items(): print("mode", k, {x: round(y, 4) if isinstance(y, float) else y for x, y in v.items()})
    for k, v in a.get("variants", {}).items(): print("variant", k, {x: round(y, 4) if isinstance(y, float) else y for x, y in v.items()})
    for k, v in d.get("partB", {}).items(): print("fit", k, v)
except Exception as e:
    print("probe parse failed", e)
PY
echo "== 3. ncu metrics of the v3 kernel (one launch)"
METRICS=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,launch__registers_per_thread
PROBE_FIT=0 PROBE_M=24576 timeout 300 ncu --metrics $METRICS --clock-control none --kernel-name regex:umma_ozaki_syrk_v3 --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r02c2_ncu_v3.csv python tools/ozaki_probe.py > /dev/null 2>&1
tail -n 12 gpurun_out/r02c2_ncu_v3.csv | cut -d, -f12- | cut -c1-170
PROBE_FIT=0 PROBE_M=24576 timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:umma_ozaki_syrk_v3 --launch-skip 3 --launch-count 1 -o gpurun_out/r02c2_v3_full -f python tools/ozaki_probe.py > /dev/null 2>&1
ls -la gpurun_out | grep r02c2
