#!/bin/bash
# zero-copy ingest flags: every GPU test, then the end-to-end leg at N=1 (and N=2 when two GPUs are visible)
source tools/r02/gpu_fn.sh
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_gpu.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/r2_tests_gpu.log
PM_BENCH_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-stages --no-cpu-baseline 2> gpurun_out/r2_ingest_n1.err > gpurun_out/r2_ingest_n1.json; grep "bench trace" gpurun_out/r2_ingest_n1.err | tail -3
python - <<'PYEOF'
import json
d=json.loads(open("gpurun_out/r2_ingest_n1.json").read().strip().splitlines()[-1])
print("N=1: value %.0f ms/step %.2f e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PYEOF
if [ $(nvidia-smi -L | wc -l) -ge 2 ]; then
PM_BENCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-configs --no-stages > gpurun_out/r2_ingest_n2.json 2> gpurun_out/r2_ingest_n2.err; grep "bench trace" gpurun_out/r2_ingest_n2.err | tail -4
python - <<'PYEOF'
import json
d=json.loads(open("gpurun_out/r2_ingest_n2.json").read().strip().splitlines()[-1])
print("N=2: value %.0f ms/step %.2f e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PYEOF
fi
