#!/bin/bash
# 2 GPUs, packed-pair ORB kernel behind the collective ingest (pm_ingest_allgather -> pack_bits_kernel incl. t4x): multi-GPU tests + ORB line
N=2
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/r2_n2pk_tests.log 2>&1; echo "multi-GPU tests exit $?"; tail -3 gpurun_out/r2_n2pk_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --kind orb --no-configs --no-stages --no-cpu-baseline > gpurun_out/r2_bench_n2_orb_packed.json 2> gpurun_out/r2_bench_n2_orb_packed.err; echo "bench N=$N orb exit $?"
python - <<PYEOF
import json
d=json.loads(open("gpurun_out/r2_bench_n2_orb_packed.json").read().strip().splitlines()[-1])
print("N=2 orb: images %s pairs %s value %.0f ms/step %.2f e2e %.0f (%.2f ms) frac %.3f" % (d["config"]["images"], d["config"]["pairs"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
PYEOF
