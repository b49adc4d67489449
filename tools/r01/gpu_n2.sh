#!/bin/bash
mkdir -p gpurun_out
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1])
    print("$2: value %.0f pairs/s ms/step %.1f | e2e %s | cfg %s" % (d["value"], d["ms_per_step"], {k: (round(v) if isinstance(v, float) else v) for k, v in d["e2e"].items()}, d["config"]["partition"]))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale2_n$N.json 2> gpurun_out/scale2_n$N.err; echo "bench N=$N exit $?"; show gpurun_out/scale2_n$N.json "N=$N sharded+async"; tail -2 gpurun_out/scale2_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 3 --warmup 3 --replicated-ingest > gpurun_out/scale2_n${N}_repl.json 2> gpurun_out/scale2_n${N}_repl.err; echo "bench N=$N replicated exit $?"; show gpurun_out/scale2_n${N}_repl.json "N=$N replicated"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale2_n${N}_ref.json 2> gpurun_out/scale2_n${N}_ref.err; echo "reference arm N=$N exit $?"; tail -c 300 gpurun_out/scale2_n${N}_ref.json
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" --timeout 300 -p no:cacheprovider 2>&1 | tail -2
