#!/bin/bash
# kind::i8 byte form: parity tests, then bench A/B (0 = i8 default, 2048 = fp16 form, 12288 = i8 64-register build)
mkdir -p gpurun_out
PY="python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x"
timeout 900 $PY -k "i8_form or fast_path or all_pairs or full_size or pair_body or cache" > gpurun_out/tests_i8.log 2>&1; echo "i8 tests exit $?"; tail -5 gpurun_out/tests_i8.log
SPECS="${SPECS:-a::0 b::12288 a2::0 b2::12288}" bash tools/gpu_i8ab.sh
