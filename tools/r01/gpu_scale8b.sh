#!/bin/bash
# 8-GPU check of the default bench (sharded upload + all-gather e2e) and of config #4's descriptor family
mkdir -p gpurun_out
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1])
    print("$2: value %.0f pairs/s ms/step %.1f | e2e %s | cfg %s | %s" % (d["value"], d["ms_per_step"], {k: (round(v) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "timing"}, d["config"]["workload"][:44], d.get("clocks")))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale3_n$N.json 2> gpurun_out/scale3_n$N.err; echo "bench N=$N exit $?"; show gpurun_out/scale3_n$N.json "N=$N"; grep -v "OMP_NUM\|^\*" gpurun_out/scale3_n$N.err | tail -2
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 --kind superpoint --images 100 --steps 2 --warmup 2 > gpurun_out/scale3_sp_n8.json 2> gpurun_out/scale3_sp_n8.err; echo "superpoint N=8 exit $?"; show gpurun_out/scale3_sp_n8.json "SP N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/scale3_ref_n8.json 2> gpurun_out/scale3_ref_n8.err; echo "reference N=8 exit $?"; tail -c 250 gpurun_out/scale3_ref_n8.json
