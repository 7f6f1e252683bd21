#!/bin/bash
# timing experiment: per-iteration cost of the partitioned PCG with parts of the communication switched off
# usage: scripts/exp_dist_breakdown.sh NGPU N_THETA N_R  (writes gpurun_out/breakdown_*.log)
N=${1:-2}; NT=${2:-2048}; NR=${3:-2048}
export FS_DIST_TIMEOUT_MS=5000
mkdir -p gpurun_out
for sk in ${SKIPS:-0 7}; do
  FS_DIST_DEBUG_SKIP=$sk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2955$sk scripts/dist_step.py --n-theta $NT --n-r $NR --steps $([ $sk = 0 ] && echo 3 || echo 0) --warmup 0 --profile-pcg 60 \
    > gpurun_out/breakdown_n${N}_skip$sk.log 2>&1
  echo "== skip $sk rc=$?"; grep "^{" gpurun_out/breakdown_n${N}_skip$sk.log | tail -1 | cut -c1-600; grep -i "error\|Traceback" -A3 gpurun_out/breakdown_n${N}_skip$sk.log | tail -8
done
timeout 300 python scripts/dist_step.py --n-theta 2048 --n-r 1024 --steps 3 --warmup 2 --profile-pcg 60 > gpurun_out/breakdown_1gpu.log 2>&1
echo "== 1 GPU rc=$?"; grep "^{" gpurun_out/breakdown_1gpu.log | tail -1 | cut -c1-600; grep -i "error\|Traceback" -A3 gpurun_out/breakdown_1gpu.log | tail -8
