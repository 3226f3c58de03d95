#!/bin/bash
# First single-GPU gpurun call of the next round: confirm the baseline, then validate and time the experimental
# variants of the tcgen05 trailing update, then ONE ncu capture with the metrics that decide between the two
# hypotheses for the 61 % main-loop ceiling (L2 -> SM feed vs smem / issue).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02_first_call.sh'
# Everything lands in gpurun_out/r02_*.  Each step has its own timeout; the experimental kernels run in subprocesses.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== 1. baseline GPU tests (ozaki + posterior + doctests)"; 
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_posterior_finitegp.py tests/test_gpu_reference_doctests.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r02_tests_baseline.log
echo "== 2. experimental variants: bit-identical to the validated kernel?"
AGP_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02_tests_experimental.log
echo "== 2b. look-ahead depth 2 (DMMA path): parity + C2 bench, default vs AGP_LOOKAHEAD=2"
AGP_LOOKAHEAD=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "logpdf_posterior or config_c2 or golden or multicolumn" 2>&1 | tail -3 | tee gpurun_out/r02_tests_lookahead2.log
timeout 200 python bench.py --steps 10 --warmup 5 --no-scaling-ref 2>/dev/null | tail -1 > gpurun_out/r02_bench_c2_la1.json
AGP_LOOKAHEAD=2 timeout 200 python bench.py --steps 10 --warmup 5 --no-scaling-ref 2>/dev/null | tail -1 > gpurun_out/r02_bench_c2_la2.json
python - <<'PY'
import json
for t in ("la1", "la2"):
    try:
        d = json.load(open("gpurun_out/r02_bench_c2_%s.json" % t)); print(t, d["value"], d["e2e"]["value"])
    except Exception as e:
        print(t, "n/a", e)
PY
echo "== 3. probe (validated modes first, variants last)"
PROBE_CLUSTER=1 timeout 120 python tools/ozaki_probe.py > gpurun_out/r02_probe_stdout.json 2> gpurun_out/r02_probe.err
cp gpurun_out/ozaki_probe.json gpurun_out/r02_ozaki_probe.json 2>/dev/null
tail -c 400 gpurun_out/r02_probe.err
echo "== 4. ncu: one launch of the validated kernel with memory-system metrics"
METRICS=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_active.avg
PROBE_FIT=0 PROBE_M=24576 timeout 300 ncu --metrics $METRICS --clock-control none --kernel-name regex:umma_ozaki_syrk_v2 --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r02_ncu_ozaki_mem.csv python tools/ozaki_probe.py > /dev/null 2>&1
tail -n 15 gpurun_out/r02_ncu_ozaki_mem.csv | cut -c1-220
echo "== 5. ncu --set full of the same launch (stall mix per PC, source view)"
PROBE_FIT=0 PROBE_M=24576 timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:umma_ozaki_syrk_v2 --launch-skip 3 --launch-count 1 -o gpurun_out/r02_ozaki_full -f python tools/ozaki_probe.py > /dev/null 2>&1
ls -la gpurun_out | tail -12
