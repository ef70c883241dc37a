#!/bin/bash
# Round-2 one-GPU check: parity suite (all GPU tests, no -x so that every failure is listed), smoke(), the default bench line
# (with its c3_512 sub-record), the reference arm, C3 / C1 / Jacobi lines, the ncu launch list of the default bench command and
# `--set full` captures of the pressure pass at 256^3 AND 512^3 plus the two advection kernels.
# Usage on the GPU box:  bash tools/gpu_r2_check.sh [tag] [quick]      (writes gpurun_out/<tag>_*)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; T=${1:-r2}; Q=${2:-full}; mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/${T}_timeline.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_gpu.txt 2>&1
stamp "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_timeline.txt
tail -15 $O/${T}_pytest_gpu.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_timeline.txt; tail -2 $O/${T}_smoke.log
stamp "bench default"
timeout 900 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; echo "rc=$?"
if [ "$Q" = "full" ]; then
stamp "bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "rc=$?"
stamp "bench C3, jacobi, C1"
timeout 300 python bench.py --workload C3 --steps 10 --no-cpu-baseline --no-extras > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
timeout 300 python bench.py --solver jacobi --no-cpu-baseline --no-extras > $O/${T}_bench_jacobi.json 2> $O/${T}_bench_jacobi.err
timeout 300 python bench.py --workload C1 --no-cpu-baseline --no-extras > $O/${T}_bench_c1.json 2> $O/${T}_bench_c1.err
fi
stamp "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/${T}_ncu_list.log 2>&1
stamp "ncu full: pressure pass 256^3, 512^3, velocity advection, density advection"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pressure -s 20 -c 1 -f -o $O/${T}_prof_pressure \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pressure -s 20 -c 1 -f -o $O/${T}_prof_pressure_c3 \
    python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras >> $O/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_advect -s 2 -c 2 -f -o $O/${T}_prof_advect \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras >> $O/${T}_ncu_full.log 2>&1
stamp "done"
for f in bench_n1 bench_reference bench_c3 bench_jacobi bench_c1; do
  [ -f "$O/${T}_$f.json" ] && python - "$O/${T}_$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline", {})
    print(sys.argv[1], "ms/step=%.3f value=%.3e e2e=%.3e launch_ms=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("launch_ms")))
    if "c3_512" in d: print("   c3_512:", {k: d["c3_512"].get(k) for k in ("ms_per_step", "value")}, d["c3_512"].get("roofline", {}).get("launch_ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
