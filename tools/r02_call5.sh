#!/bin/bash
# Round 2, multi-GPU call:  gpurun --gpus N -- 'bash tools/r02_call5.sh N'
# (1) correctness of the distributed fit + posterior handle under the pipelined schedule (default) and the plain one,
# (2) C4 bench lines for the schedules / reserve sizes.  Never under ncu.
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run; export N
echo "== dist_fit_check, pipelined schedule (default)"
timeout 600 bash -c "run 29551 tests/dist_fit_check.py" > gpurun_out/r02c5_check_sched2_${N}.full.log 2>&1
grep "^\[rank\|DIST_" gpurun_out/r02c5_check_sched2_${N}.full.log | cut -c1-330 | tee gpurun_out/r02c5_check_sched2_${N}.log
echo "== dist_fit_check stress with forced 512-wide tcgen05 panels at small n (few blocks per rank)"
AGP_NB=512 AGP_FP64_MODE=1 DIST_CHECK_ONLY_STRESS=1 DIST_CHECK_SIZES=${STRESS_SIZES:-2048,2560,3072,4096,5120} timeout 600 bash -c "run 29553 tests/dist_fit_check.py" > gpurun_out/r02c5_check_stress_${N}.full.log 2>&1
grep "^\[rank\|DIST_" gpurun_out/r02c5_check_stress_${N}.full.log | cut -c1-330 | tee gpurun_out/r02c5_check_stress_${N}.log
if [ "${CHECK_PLAIN:-1}" = "1" ]; then
echo "== dist_fit_check, plain look-ahead schedule"
AGP_DIST_SCHED=0 timeout 600 bash -c "run 29552 tests/dist_fit_check.py" 2>&1 | grep -v "^W\|OMP_NUM" | tail -4 | tee gpurun_out/r02c5_check_sched0_${N}.log
fi
# quick schedule sweep: "sched reserve defer chunk"
IFS=";" read -ra SWEEP <<< "${SWEEP:-}"
port=29580
for cfg in "${SWEEP[@]}"; do
  set -- $cfg
  port=$((port + 1))
  echo "== quick C4 N=$N sched=$1 reserve=$2 defer=$3 chunk=$4"
  AGP_DIST_SCHED=$1 AGP_DIST_RESERVE_SMS=$2 AGP_DIST_DEFER=$3 AGP_OZAKI_CHUNK=$4 timeout 600 bash -c "run $port bench.py --gpus $N --steps 3 --warmup 3 --quick" 2>/dev/null | tail -1 | cut -c1-400
done
port=29560
IFS=";" read -ra CFG_LIST <<< "${CFGS:-2 16;0 16;2 8;2 32}"
for cfg in "${CFG_LIST[@]}"; do
  [ "${RUN_BENCH:-1}" = "1" ] || break
  set -- $cfg
  port=$((port + 1))
  echo "== bench C4 N=$N AGP_DIST_SCHED=$1 reserve=$2"
  AGP_DIST_SCHED=$1 AGP_DIST_RESERVE_SMS=$2 timeout 900 bash -c "run $port bench.py --gpus $N --steps 3 --warmup 3" 2>/dev/null | tail -1 > gpurun_out/r02c5_bench_c4_${N}gpu_sched$1_res$2.json
  python - "$N" "$1" "$2" <<'PY'
import json, sys
n, sc, rs = sys.argv[1:4]
try:
    d = json.loads(open("gpurun_out/r02c5_bench_c4_%sgpu_sched%s_res%s.json" % (n, sc, rs)).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "phases", {k: round(v, 1) for k, v in d["phases_ms"].items()}, "logpdf", d["result"],
          "parity", d["parity"] and d["parity"].get("ok"), "trailing_ms", round(r["kernel_ms_per_step"], 1), "frac", r["frac"])
except Exception as e:
    print("parse failed", e)
PY
done
if [ "${RUN_C5:-0}" = "1" ]; then
  echo "== bench C5 (VFE, N = 10^6, M = 8192, fp32) on $N GPUs"
  timeout 900 bash -c "run 29599 bench.py --gpus $N --workload C5 --steps 3 --warmup 3" 2>/dev/null | tail -1 > gpurun_out/r02c5_bench_c5_${N}gpu.json
  python - "$N" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02c5_bench_c5_%sgpu.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("C5 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "phases", {k: round(v, 1) for k, v in d["phases_ms"].items()}, "elbo", d["result"], "parity", d["parity"], "roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac")})
except Exception as e:
    print("parse failed", e)
PY
fi
