#!/bin/bash
# 1 -> 8 GPU scaling of the default bench on ONE box (run under `gpurun --gpus 8`)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1])
    print("$2: value %.0f pairs/s ms/step %.1f | e2e %s | cfg %s | clocks %s" % (d["value"], d["ms_per_step"], {k: (round(v) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "timing"}, d["config"]["workload"][:40], d.get("clocks")))
except Exception as e: print("parse fail", e)
PYEOF
}
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "bench N=$N exit $?"; show gpurun_out/scale_n$N.json "N=$N"; tail -2 gpurun_out/scale_n$N.err
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "bench N=1 exit $?"; show gpurun_out/scale_n1.json "N=1"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --kind superpoint --images 64 --steps 2 --warmup 2 > gpurun_out/scale_sp_n8.json 2> gpurun_out/scale_sp_n8.err; echo "superpoint N=8 exit $?"; show gpurun_out/scale_sp_n8.json "SP N=8"
