#!/bin/bash
# Round 2, first single-GPU call: (1) in-situ parity of the tcgen05 path + the whole default GPU suite, (2) the opt-in
# tests of everything written but never run on a device, (3) probe of the trailing kernel's phases and variants,
# (4) ncu: memory metrics + full set of the trailing kernel, per-kernel captures of gram / solves / panel kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
echo "== 1a. new in-situ tcgen05 parity tests"
timeout 600 python -m pytest tests/test_gpu_tcgen05_insitu.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_t_insitu.log
echo "== 1b. default GPU suite"
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_tcgen05_insitu.py 2>&1 | tail -6 | tee gpurun_out/r02_t_default.log
echo "== 2. opt-in (never run on a device) tests"
AGP_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02_t_experimental.log
echo "== 2b. look-ahead 2 on the DMMA path"
AGP_LOOKAHEAD=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "logpdf_posterior or config_c2 or golden or multicolumn" 2>&1 | tail -3 | tee gpurun_out/r02_t_la2.log
for la in 1 2; do AGP_LOOKAHEAD=$la timeout 120 python tools/fit_once.py 4096 8 8 | tee gpurun_out/r02_c2_la$la.txt; done
echo "== 3. probe"
PROBE_CLUSTER=1 PROBE_FIT=0 timeout 200 python tools/ozaki_probe.py > gpurun_out/r02_probe_stdout.json 2> gpurun_out/r02_probe.err
cp gpurun_out/ozaki_probe.json gpurun_out/r02_ozaki_probe.json 2>/dev/null
tail -c 600 gpurun_out/r02_probe.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_ozaki_probe.json"))["partA"]
    print("fixed_ms", d["fixed_ms"])
    for k, v in d["modes"].items(): print("mode", k, {a: round(b, 4) if isinstance(b, float) else b for a, b in v.items()})
    for k, v in d.get("variants", {}).items(): print("variant", k, v)
except Exception as e:
    print("probe parse failed", e)
PY
echo "== 4a. ncu memory metrics, validated trailing kernel"
METRICS=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max
PROBE_FIT=0 PROBE_M=24576 timeout 300 ncu --metrics $METRICS --clock-control none --kernel-name regex:umma_ozaki_syrk_v2 --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r02_ncu_ozaki_mem.csv python tools/ozaki_probe.py > /dev/null 2>&1
tail -n 16 gpurun_out/r02_ncu_ozaki_mem.csv | cut -d, -f12- | cut -c1-160
echo "== 4b. ncu --set full, same launch"
PROBE_FIT=0 PROBE_M=24576 timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:umma_ozaki_syrk_v2 --launch-skip 3 --launch-count 1 -o gpurun_out/r02_ozaki_full -f python tools/ozaki_probe.py > /dev/null 2>&1
echo "== 4c. per-kernel captures at C2 (N=4096 D=8) and C4q (N=16384 D=64)"
KM=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 300 ncu --metrics $KM --clock-control none -k regex:'gram_kernel|bwd_solve_kernel|trsm_sub_f64|potrf_factor_only_f64|trtri_strips|extract_v|border_init|finalize' --launch-skip 0 --launch-count 60 --csv --log-file gpurun_out/r02_ncu_kernels_c2.csv python tools/fit_once.py 4096 8 1 > /dev/null 2>&1
timeout 300 ncu --metrics $KM --clock-control none -k regex:'gram_kernel|bwd_solve_kernel|ozaki_slice|ozaki_rowscale' --launch-skip 0 --launch-count 12 --csv --log-file gpurun_out/r02_ncu_kernels_c4q.csv python tools/fit_once.py 16384 64 1 > /dev/null 2>&1
echo "== 4d. launch list of one C4h fit (shares)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_c4h.csv python tools/fit_once.py 32768 64 1 > /dev/null 2>&1
ls -la gpurun_out | tail -20
