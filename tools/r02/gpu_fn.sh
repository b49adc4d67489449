# helper for the round-2 GPU scripts: run <tag> <bench args...> -> gpurun_out/r2_<tag>.json + a one-line summary
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2_$tag.json 2> gpurun_out/r2_$tag.err; echo "$tag exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/r2_$tag.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("$tag: value %.0f pairs/s ms/step %.2f | knn %.0f %s frac %.3f avg %.3f ms share %s | e2e %s | clocks %s" % (d["value"], d["ms_per_step"], r["achieved"], r["unit"], r["frac"], r["avg_launch_ms"], r["share_of_step"], e.get("value"), d["clocks"]))
    for k,c in (d.get("configs") or {}).items():
        rr=c["roofline"]; ee=c.get("e2e") or {}
        print("   %s: value %.0f ms/step %.1f | frac %.3f avg %.3f share %s | e2e %s | setup %ss | %s" % (k, c["value"], c["ms_per_step"], rr["frac"], rr["avg_launch_ms"], rr["share_of_step"], ee.get("value"), c["setup_s"], c["clocks"]))
    if d.get("stages") and d["stages"].get("ransac_heavy"): print("   ransac_heavy:", d["stages"]["ransac_heavy"]["value"])
except Exception as ex: print("parse fail", ex); print(open("gpurun_out/r2_$tag.err").read()[-1500:])
PYEOF
}
