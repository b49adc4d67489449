#!/bin/bash
# 2 GPUs: host / device timeline of one end-to-end step (PM_TRACE + PM_BENCH_TRACE)
N=2
PM_BENCH_TRACE=1 PM_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2 --warmup 1 --no-configs --no-stages > gpurun_out/r2_trace_n2.json 2> gpurun_out/r2_trace_n2.err; echo "exit $?"
grep -E "bench trace" gpurun_out/r2_trace_n2.err | tail -8
grep -E "pm trace" gpurun_out/r2_trace_n2.err | tail -120 | cut -c1-200 > gpurun_out/r2_trace_n2_tail.txt; wc -l gpurun_out/r2_trace_n2_tail.txt
PM_BENCH_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-configs --no-stages --no-cpu-baseline 2>&1 >/dev/null | grep "bench trace" | tail -3
