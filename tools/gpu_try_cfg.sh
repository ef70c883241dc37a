#!/bin/bash
# A/B of pressure-kernel configurations on one GPU: parity subset + bench lines per SMK_FUSED_CFG value.
# usage: bash tools/gpu_try_cfg.sh "0 20" ["C2 C3"]
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
CFGS=${1:-"0 20"}; WLS=${2:-"C2 C3"}
for c in $CFGS; do
  echo "== SMK_FUSED_CFG=$c: parity subset"
  SMK_FUSED_CFG=$c timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_slab_gpu.py -m gpu -x -q -k "fused or c1 or c2 or slab or peer or balanced" 2>&1 | tail -3
  for w in $WLS; do
    steps=20; [ "$w" = "C3" ] && steps=6
    SMK_FUSED_CFG=$c timeout 300 python bench.py --workload $w --steps $steps --no-cpu-baseline > $O/try_${w}_cfg$c.json 2> $O/try_${w}_cfg$c.err
    python - "$O/try_${w}_cfg$c.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "ms/step=%.3f value=%.3e launch_ms=%.4f stages=%s" % (d["ms_per_step"], d["value"], r["launch_ms"], {k: round(v,3) for k,v in r["stage_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
