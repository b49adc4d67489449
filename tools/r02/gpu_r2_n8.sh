#!/bin/bash
# 8 GPUs: the driver's scaling command (default bench line incl. the strong-scaling configs), once
N=${1:-8}
nvidia-smi -L | wc -l
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N exit $? in $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N main: value %.0f ms/step %.2f e2e %.0f (%.2f ms) ag %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("allgather_bytes_per_step")))
    for k,c in d["configs"].items():
        print(k, "value %.0f ms/step %.1f e2e %.0f (%.1f ms) setup %s" % (c["value"], c["ms_per_step"], c["e2e"]["value"], c["e2e"]["ms_per_step"], c["setup_s"]), c["clocks"])
    print("heavy", d["stages"]["ransac_heavy"]["value"] if d.get("stages") else None)
except Exception as ex: print("parse fail", ex)
PYEOF
