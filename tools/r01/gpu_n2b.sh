#!/bin/bash
# 2-GPU sanity + numbers after the copy-kernel change: sharded ingest (NCCL all-gather) and the multi-device handle test
mkdir -p gpurun_out
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1])
    print("$2: value %.0f pairs/s ms/step %.1f | e2e %s | cfg %s" % (d["value"], d["ms_per_step"], {k: (round(v) if isinstance(v, float) else v) for k, v in d["e2e"].items()}, d["config"]["partition"]))
except Exception as e: print("$2 parse fail", e)
PYEOF
}
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale4_n$N.json 2> gpurun_out/scale4_n$N.err; echo "bench N=$N exit $?"; show gpurun_out/scale4_n$N.json "N=$N sift"; tail -2 gpurun_out/scale4_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --kind orb --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/scale4_orb_n$N.json 2> gpurun_out/scale4_orb_n$N.err; echo "bench orb N=$N exit $?"; show gpurun_out/scale4_orb_n$N.json "N=$N orb"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py -q -k "multi_device or shim" --timeout 300 -p no:cacheprovider 2>&1 | tail -2
