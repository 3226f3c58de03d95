#!/bin/bash
# First multi-GPU call of the next round (N = 2, then the same with --gpus 8 when a large box answers):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/r02_dist_call.sh 2'
# 1. correctness of the default and of the experimental owner-first schedule (tests/dist_fit_check.py prints DIST_OK),
# 2. C4 bench lines for both schedules and two reserve sizes.  Never under ncu.
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== default schedule"; timeout 600 bash -c "$(declare -f run); N=$N; run 29551 tests/dist_fit_check.py" 2>&1 | tail -3 | tee gpurun_out/r02_dist_check_default.log
echo "== AGP_DIST_SCHED=1";   AGP_DIST_SCHED=1 timeout 600 bash -c "$(declare -f run); N=$N; run 29552 tests/dist_fit_check.py" 2>&1 | tail -3 | tee gpurun_out/r02_dist_check_sched1.log
echo "== AGP_OZAKI_GROUPED=1"; AGP_OZAKI_GROUPED=1 timeout 600 bash -c "$(declare -f run); N=$N; run 29553 tests/dist_fit_check.py" 2>&1 | tail -3 | tee gpurun_out/r02_dist_check_grouped.log
echo "== both";                AGP_OZAKI_GROUPED=1 AGP_DIST_SCHED=1 timeout 600 bash -c "$(declare -f run); N=$N; run 29554 tests/dist_fit_check.py" 2>&1 | tail -3 | tee gpurun_out/r02_dist_check_both.log
port=29560
for cfg in "0 16 0" "0 16 1" "1 16 0" "1 8 1" "1 16 1" "1 24 1"; do
  set -- $cfg
  port=$((port + 1))
  echo "== bench C4 N=$N AGP_DIST_SCHED=$1 reserve=$2 AGP_OZAKI_GROUPED=$3"
  AGP_OZAKI_GROUPED=$3 AGP_DIST_SCHED=$1 AGP_DIST_RESERVE_SMS=$2 timeout 600 bash -c "$(declare -f run); N=$N; run $port bench.py --gpus $N --steps 3 --warmup 3" 2>/dev/null | tail -1 | tee gpurun_out/r02_bench_c4_${N}gpu_sched$1_res$2_grp$3.json | cut -c1-300
done
