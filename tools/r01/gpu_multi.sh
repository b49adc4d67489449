#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" --timeout 300 -p no:cacheprovider > gpurun_out/tests_multi.log 2>&1; echo "multi-device test exit $?"; tail -3 gpurun_out/tests_multi.log
show() { python - <<PYEOF
import json
try:
    d=json.loads([l for l in open("$1").read().strip().splitlines() if l.startswith("{")][-1])
    print("$2: value %.0f pairs/s ms/step %.1f | e2e %s | cfg %s" % (d["value"], d["ms_per_step"], {k: (round(v) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "timing"}, d["config"]["workload"][:40]))
except Exception as e: print("parse fail", e)
PYEOF
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; show gpurun_out/bench_n$N.json "N=$N"; tail -3 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --sharded-ingest > gpurun_out/bench_n${N}_sharded.json 2> gpurun_out/bench_n${N}_sharded.err; echo "bench N=$N sharded exit $?"; show gpurun_out/bench_n${N}_sharded.json "N=$N sharded"; tail -3 gpurun_out/bench_n${N}_sharded.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_n${N}_ref.json 2> gpurun_out/bench_n${N}_ref.err; echo "reference arm N=$N exit $?"; tail -c 400 gpurun_out/bench_n${N}_ref.json
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --cpu-seconds 6 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench N=1 exit $?"; show gpurun_out/bench_n1.json "N=1"; python -c "
import json;d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]);print('stages',d['stages']);print('cpu',d['cpu_baseline'])"
