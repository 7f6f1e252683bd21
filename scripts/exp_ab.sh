#!/bin/bash
# A/B of environment switches on the N=1 bench: scripts/exp_ab.sh "FS_X=0 FS_Y=1" "FS_X=1" ...   (one quoted group per run)
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - "$cfg" gpurun_out/ab_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    it = r.get("us_per_pcg_iteration", {})
    print(f"{sys.argv[1]:55s} value {d['value']:.2f} ms/step {d['ms_per_step']:.3f} | A*p {it.get('A*p',0):.1f} V {it.get('V-cycle',0):.1f} vec {it.get('vector ops + dots',0):.1f} tot {it.get('total',0):.1f} | top {r['us_per_launch']:.1f} us frac {r['frac']:.3f} | its {sum(map(sum, d['config']['cg_iters_per_step']))}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
