#!/bin/bash
# 2 GPUs, block-cyclic shares: the bench line at N=2 (torchrun) -- headline config only (weak scaling) + multi-GPU tests
source tools/r02/gpu_fn.sh
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_tests_multi.log 2>&1; echo "multi tests exit $?"; tail -3 gpurun_out/r2_tests_multi.log
N=2
for rep in a b; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-configs --no-stages > gpurun_out/r2_bench_n${N}_cyc_$rep.json 2> gpurun_out/r2_bench_n${N}_cyc_$rep.err; echo "bench N=$N exit $?"
python - <<PYEOF
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_cyc_$rep.json").read().strip().splitlines()[-1])
print("N=$N main: value %.0f ms/step %.2f e2e %.0f (%.2f ms) ag %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("allgather_bytes_per_step")))
PYEOF
done
timeout 900 python bench.py --no-configs --no-stages --no-cpu-baseline > gpurun_out/r2_bench_n1_same_box.json 2>/dev/null
python - <<PYEOF
import json
d=json.loads(open("gpurun_out/r2_bench_n1_same_box.json").read().strip().splitlines()[-1])
print("N=1 main: value %.0f ms/step %.2f e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PYEOF
