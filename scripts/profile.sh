#!/bin/bash
# ncu evidence for profiles/: run on the GPU box via
#   gpurun --timeout 1500 -- 'bash scripts/profile.sh r01'
# 1) launch list of one default bench step (AMG-PCG): per-launch device time, cold cache, serialised
# 2) full-set capture of the dominant kernel of that step (fine-level k_spmv_warp<EPI_AX, DOT> = the CG A*p)
# 3) full-set capture of the persistent Jacobi-CG kernel (200 iterations, 4M triangles)
# 4) DRAM bytes of the persistent kernel with caches left alone (steady-state L2 residency)
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
export PYTHONPATH=.
python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > $out/${tag}_plain_bench.json 2> $out/${tag}_plain_bench.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > $out/${tag}_ncu_launches.log 2>&1
python scripts/prof_amg.py 4 > $out/${tag}_plain_amg.log 2>&1 || { echo "plain amg failed"; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_spmv_warp<.int.0, .bool.1" -s 3 -c 1 -o $out/${tag}_spmv_warp \
    python scripts/prof_amg.py 4 > $out/${tag}_ncu_spmv.log 2>&1
python scripts/prof_cg.py 200 > $out/${tag}_plain_cg.log 2>&1 || { echo "plain cg failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -c 1 -o $out/${tag}_cg_persistent \
    python scripts/prof_cg.py 200 > $out/${tag}_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none \
    -k regex:k_cg_persistent -c 1 --csv --log-file $out/${tag}_cg_dram_nocachectl.csv \
    python scripts/prof_cg.py 200 > $out/${tag}_ncu_dram.log 2>&1
tail -2 $out/${tag}_ncu_full.log
