mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 900 python bench.py "$@" --no-cpu-baseline > gpurun_out/misc_$tag.json 2> gpurun_out/misc_$tag.err; echo "$tag exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/misc_$tag.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$tag: value %.0f pairs/s ms/step %.1f | knn %.2f %s frac %.3f avg %.3f ms share %.3f | e2e %.0f | put %d inl %d | %s" % (d["value"], d["ms_per_step"], r["achieved"], r["unit"], r["frac"], r["avg_launch_ms"], r["share_of_step"], d["e2e"]["value"], d["putative_matches_per_step"], d["inliers_per_step"], d["clocks"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/misc_$tag.err").read()[-1500:])
PYEOF
}
