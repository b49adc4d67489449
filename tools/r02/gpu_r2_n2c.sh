#!/bin/bash
# 2 GPUs: head-first collective ingest (default) against the two-phase order (debug bit 24); multi-GPU tests
source tools/r02/gpu_fn.sh
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_tests_multi.log 2>&1; echo "multi tests exit $?"; tail -3 gpurun_out/r2_tests_multi.log
N=2
for v in head:0 twophase:16777216 head2:0 sp:0; do
name=${v%%:*}; flags=${v##*:}
extra=""; if [ $name = sp ]; then extra="--kind superpoint --images 100"; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-configs --no-stages --debug-flags $flags $extra > gpurun_out/r2_bench_n${N}_$name.json 2> gpurun_out/r2_bench_n${N}_$name.err; echo "bench N=$N $name exit $?"
python - <<PYEOF
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_$name.json").read().strip().splitlines()[-1])
print("N=$N $name: value %.0f ms/step %.2f e2e %.0f (%.2f ms) ag %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("allgather_bytes_per_step")))
PYEOF
done
