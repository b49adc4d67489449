#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" --timeout 300 -p no:cacheprovider > gpurun_out/tests_multi.log 2>&1; echo "multi-device test exit $?"; tail -3 gpurun_out/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; tail -c 2500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench N=1 exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]);print('N=1 value',d['value'],'e2e',d['e2e']['value'])"
