#!/bin/bash
# Round 2, call 9 (1 GPU): final regression on the defaults + the records that go under profiles/
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02c9_tests.log
echo "== 2. smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== 3. 1-rank distributed path, host vs device pointers (allocation order fix)"
timeout 300 python tools/dist1_probe.py 32768 64 3 2>&1 | tail -12 | tee gpurun_out/r02c9_dist1.log
echo "== 4. bench: C4 default line, C3, C5 at full size on one GPU"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c9_bench_c4_1gpu.json 2> gpurun_out/r02c9_bench_c4.err; tail -c 200 gpurun_out/r02c9_bench_c4.err
timeout 600 python bench.py --workload C3 --steps 5 --warmup 3 > gpurun_out/r02c9_bench_c3.json 2> gpurun_out/r02c9_bench_c3.err; tail -c 200 gpurun_out/r02c9_bench_c3.err
timeout 900 python bench.py --workload C5 --steps 3 --warmup 3 > gpurun_out/r02c9_bench_c5_1gpu.json 2> gpurun_out/r02c9_bench_c5.err; tail -c 200 gpurun_out/r02c9_bench_c5.err
python - <<'PY'
import json
for t in ("c4_1gpu", "c3", "c5_1gpu"):
    try:
        d = json.loads(open("gpurun_out/r02c9_bench_%s.json" % t).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(t, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "phases", {k: round(v, 2) for k, v in d["phases_ms"].items()},
              "parity", d["parity"] and (d["parity"].get("ok"), d["parity"].get("rel_err")), "roofline", {k: r.get(k) for k in ("achieved", "peak", "frac", "kernel_ms_per_step", "frac_of_nominal_4500")},
              "clocks", d["clocks"], "c2", d.get("c2") and (d["c2"]["value"], d["c2"]["e2e"]))
    except Exception as e:
        print(t, "parse failed", e)
PY
echo "== 5. ncu: launch list of one C4 fit + full capture of its largest trailing update"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02c9_launches_c4.csv python tools/fit_once.py 65536 64 1 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:umma_ozaki_syrk_v3 --launch-skip 1 --launch-count 1 -o gpurun_out/r02c9_v3_c4_full -f python tools/fit_once.py 65536 64 1 > /dev/null 2>&1
ls -la gpurun_out | grep r02c9
