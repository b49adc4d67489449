#!/bin/bash
# 2 GPUs: order of the collective ingest -- progressive chunks (default) / head + rest (bit 25) / two-phase (bit 24)
N=2
for v in prog:0 head:33554432 twophase:16777216 prog2:0; do
name=${v%%:*}; flags=${v##*:}
PM_BENCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-configs --no-stages --debug-flags $flags > gpurun_out/r2_bench_n${N}_$name.json 2> gpurun_out/r2_bench_n${N}_$name.err; echo "bench N=$N $name exit $?"
grep "bench trace" gpurun_out/r2_bench_n${N}_$name.err | tail -2
python - <<PYEOF
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_$name.json").read().strip().splitlines()[-1])
print("N=$N $name: value %.0f ms/step %.2f e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PYEOF
done
