#!/bin/bash
# ncu evidence for profiles/: run on the GPU box via
#   gpurun --timeout 1500 -- 'bash scripts/profile.sh r01'            (all four steps)
#   gpurun --timeout 1500 -- 'STEPS="1 2" bash scripts/profile.sh r01' (a subset)
# 1) launch list of one default bench step (AMG-PCG): per-launch device time, cold cache, serialised
# 2) full-set capture of the dominant kernel of that step, k_spmv_sell, in its two big instances: the finest
#    up-sweep of the folded V-cycle (split form, fp32 values, fused r.z) and the CG's A*p (fp64 values, fused p.Ap)
# 3) full-set capture of the persistent Jacobi-CG kernel (200 iterations, 4M triangles)
# 4) DRAM bytes of the persistent kernel with caches left alone (steady-state L2 residency)
# Every ncu pass is preceded by the same command without ncu (must exit 0).
set -u
tag=${1:-r01}
steps=${STEPS:-"1 2 3 4"}
out=gpurun_out
mkdir -p $out
export PYTHONPATH=.
has() { [[ " $steps " == *" $1 "* ]]; }
if has 1; then
  python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > $out/${tag}_plain_bench.json 2> $out/${tag}_plain_bench.err || { echo "plain bench failed"; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > $out/${tag}_ncu_launches.log 2>&1
fi
if has 2; then
  python scripts/prof_amg.py 6 > $out/${tag}_plain_amg.log 2>&1 || { echo "plain amg failed"; exit 1; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_spmv_sell<.bool.(0|1), .bool.1" --launch-skip 6 -c 2 \
      -o $out/${tag}_spmv_sell -f python scripts/prof_amg.py 6 > $out/${tag}_ncu_spmv.log 2>&1
  tail -2 $out/${tag}_ncu_spmv.log
fi
if has 3; then
  python scripts/prof_cg.py 200 > $out/${tag}_plain_cg.log 2>&1 || { echo "plain cg failed"; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -c 1 -o $out/${tag}_cg_persistent -f \
      python scripts/prof_cg.py 200 > $out/${tag}_ncu_full.log 2>&1
  tail -2 $out/${tag}_ncu_full.log
fi
if has 4; then
  python scripts/prof_cg.py 200 > $out/${tag}_plain_cg.log 2>&1 || { echo "plain cg failed"; exit 1; }
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none \
      -k regex:k_cg_persistent -c 1 --csv --log-file $out/${tag}_cg_dram_nocachectl.csv \
      python scripts/prof_cg.py 200 > $out/${tag}_ncu_dram.log 2>&1
fi
if has 5; then   # roofline records of the rest of the path: DRAM bytes and duration per launch (4M triangles, 4M tracers)
  python scripts/prof_path.py > $out/${tag}_plain_path.log 2>&1 || { echo "plain path run failed"; tail -5 $out/${tag}_plain_path.log; exit 1; }
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
      --clock-control none --kernel-name-base demangled \
      -k regex:"k_element_stiffness|k_assemble|k_div_elem|k_div_node|k_grad_elem|k_grad_node|k_locate|k_advect|k_tracer_step|k_backtrace" \
      -c 60 --csv --log-file $out/${tag}_path_kernels.csv python scripts/prof_path.py > $out/${tag}_ncu_path.log 2>&1
  tail -2 $out/${tag}_ncu_path.log
fi
if has 6; then   # A/B: the bulk-async (TMA) staged SELL kernel against the register-staged one (same launches)
  FS_SELL_TMA=2 python scripts/prof_amg.py 6 > $out/${tag}_plain_amg_tma.log 2>&1 || { echo "plain amg (tma) failed"; exit 1; }
  FS_SELL_TMA=2 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_spmv_sell_bulk<.bool.(0|1), .bool.1" --launch-skip 6 -c 2 \
      -o $out/${tag}_spmv_sell_bulk -f python scripts/prof_amg.py 6 > $out/${tag}_ncu_spmv_tma.log 2>&1
  tail -2 $out/${tag}_ncu_spmv_tma.log
fi
if has 7; then   # the five k_spmv_sell launches of one AMG-PCG iteration (A*p, finest / level-1 restriction, level-1 / finest up-sweep)
  python scripts/prof_amg.py 6 > $out/${tag}_plain_amg_pk.log 2>&1 || { echo "plain amg (packed) failed"; exit 1; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_spmv_sell<" --launch-skip 9 -c 5 \
      -o $out/${tag}_spmv_pk -f python scripts/prof_amg.py 6 > $out/${tag}_ncu_spmv_pk.log 2>&1
  tail -2 $out/${tag}_ncu_spmv_pk.log
fi
echo "profile.sh done: $steps"
