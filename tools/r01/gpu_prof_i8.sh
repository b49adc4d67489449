#!/bin/bash
# Round-end evidence for the kind::i8 SIFT path: smoke, default bench (both arms), launch list, ncu --set full of
# the tensor kernel and of the fix-up, plus bench lines of the other configs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" > gpurun_out/cpu_info.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sift_i8.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
CMD2="python bench.py --images 23 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stages"
$CMD2 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:l2_i8x2 -s 1 -c 1 -f -o gpurun_out/prof_i8x2 $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full i8x2 exit $?"; tail -2 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:l2_fixup_i8 -s 1 -c 1 -f -o gpurun_out/prof_fixup_i8 $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full fixup exit $?"; tail -2 gpurun_out/ncu_full2.log
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
    print("$2: value %.0f pairs/s ms/step %.1f | knn %.3f ms frac %.3f share %.2f | e2e %.0f | %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["share_of_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
timeout 600 python bench.py --outlier-frac 0.5 --steps 3 --warmup 2 --no-cpu-baseline --no-stages > gpurun_out/bench_sift_out50.json 2>gpurun_out/bench_sift_out50.err; show gpurun_out/bench_sift_out50.json "sift outlier 0.5"
timeout 900 python bench.py --kind orb --images 500 --steps 2 --warmup 1 --no-cpu-baseline --no-stages > gpurun_out/bench_orb500.json 2>gpurun_out/bench_orb500.err; show gpurun_out/bench_orb500.json "orb 500 (config 3)"
timeout 900 python bench.py --kind superpoint --images 64 --steps 3 --warmup 2 --no-cpu-baseline --no-stages > gpurun_out/bench_sp64.json 2>gpurun_out/bench_sp64.err; show gpurun_out/bench_sp64.json "superpoint 64"
for kp in 4096 16384; do
  timeout 900 python bench.py --kind sift --images 64 --kp $kp --steps 3 --warmup 2 --no-cpu-baseline --no-stages > gpurun_out/bench_sift_kp$kp.json 2>gpurun_out/bench_sift_kp$kp.err; show gpurun_out/bench_sift_kp$kp.json "sift 64 x $kp"
  timeout 900 python bench.py --kind orb --images 64 --kp $kp --steps 3 --warmup 2 --no-cpu-baseline --no-stages > gpurun_out/bench_orb_kp$kp.json 2>gpurun_out/bench_orb_kp$kp.err; show gpurun_out/bench_orb_kp$kp.json "orb 64 x $kp"
done
