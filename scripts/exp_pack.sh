#!/bin/bash
# packed SELL entries: parity subset, then A/B on the N=1 bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "packed_sell or cycle_forms or synthetic_amg or one_block or amg_pcg_matches" 2>&1 | tail -15 > gpurun_out/pack_tests.log
tail -3 gpurun_out/pack_tests.log
bash scripts/exp_ab.sh "$@" 2>&1 | tee gpurun_out/pack_ab.log
