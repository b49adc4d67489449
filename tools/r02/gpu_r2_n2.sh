#!/bin/bash
# 2 GPUs: multi-GPU tests, then the bench line at N=2 (torchrun), then N=1 on the same box for the scaling ratio
source tools/r02/gpu_fn.sh
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "multi or collective" > gpurun_out/r2_tests_multi.log 2>&1; echo "multi tests exit $?"; tail -8 gpurun_out/r2_tests_multi.log
N=${1:-2}
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
except Exception as ex: print("parse fail", ex)
PYEOF
