#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; T=${1:-ab}; mkdir -p $O
export SMK_PASS_KERNEL=tma
timeout 900 python -m pytest tests -m gpu -q -k "not multi_gpu" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${T}_pytest.log | cut -c1-200
for w in C2 C3; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_tma_${w}.json 2> $O/${T}_tma_${w}.err
    python - "$O/${T}_tma_${w}.json" tma $w <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], sys.argv[3], "ms/step=%.3f pass_ms=%.4f frac_compulsory=%.3f stages=%s" % (d["ms_per_step"], r["launch_ms"], r["frac_compulsory"], {k: round(v,3) for k,v in r["stage_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
SMK_PASS_DEBUG=1 python tools/cta_times.py C2; SMK_PASS_DEBUG=1 python tools/cta_times.py C3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pressure_tma -s 20 -c 1 -f -o $O/${T}_prof_tma \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/${T}_ncu.log 2>&1
