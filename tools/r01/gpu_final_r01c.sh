#!/bin/bash
# Round-1 closing evidence (final build): smoke, both bench arms on the default config, config #3 (ORB-500), SuperPoint-100, ORB-100,
# launch list of the default bench.  (Every GPU test ran green on this build in tools/gpu_scanfix.sh.)
mkdir -p gpurun_out
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
    print("$2: value %.0f pairs/s ms/step %.1f | knn %.3f ms frac %.3f share %.2f | e2e %.0f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["share_of_step"], d["e2e"]["value"] if d.get("e2e") else -1, d["clocks"]))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; show gpurun_out/bench_default.json "default (sift 100)"
timeout 900 python bench.py --kind orb --images 500 --steps 2 --warmup 1 --no-cpu-baseline --no-stages > gpurun_out/bench_orb500.json 2>gpurun_out/bench_orb500.err; show gpurun_out/bench_orb500.json "orb 500 (config 3)"
timeout 900 python bench.py --kind orb --images 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_orb100.json 2>gpurun_out/bench_orb100.err; show gpurun_out/bench_orb100.json "orb 100"
timeout 900 python bench.py --kind superpoint --images 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_sp100.json 2>gpurun_out/bench_sp100.err; show gpurun_out/bench_sp100.json "superpoint 100"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_sift_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches sift exit $?"
