#!/bin/bash
# One-GPU check of the balanced pass schedule: parity suite, A/B bench lines, ncu launch list + one full capture.
# Usage on the GPU box: bash tools/gpu_check_balanced.sh   (writes gpurun_out/*)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.txt; }
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/timeline.txt
tail -5 $O/pytest_gpu.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/timeline.txt; tail -2 $O/smoke.log
stamp "bench default"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"
stamp "bench grid (SMK_PASS_BALANCED=0)"
SMK_PASS_BALANCED=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1_grid.json 2> $O/bench_n1_grid.err
stamp "bench C3"
timeout 300 python bench.py --workload C3 --steps 6 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
SMK_PASS_BALANCED=0 timeout 300 python bench.py --workload C3 --steps 6 --no-cpu-baseline > $O/bench_c3_grid.json 2> $O/bench_c3_grid.err
stamp "bench jacobi"
timeout 300 python bench.py --solver jacobi --no-cpu-baseline > $O/bench_jacobi.json 2> $O/bench_jacobi.err
SMK_PASS_BALANCED=0 timeout 300 python bench.py --solver jacobi --no-cpu-baseline > $O/bench_jacobi_grid.json 2> $O/bench_jacobi_grid.err
stamp "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_list.log 2>&1
stamp "ncu full, one balanced pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pressure_reg_bal -s 20 -c 1 -f -o $O/prof_bal \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
stamp "done"
for f in bench_n1 bench_n1_grid bench_c3 bench_c3_grid bench_jacobi bench_jacobi_grid; do
  python - "$O/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], "ms/step=%.3f value=%.3e e2e=%.3e launch_ms=%.4f kernel=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"], r["kernel"][:40]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
