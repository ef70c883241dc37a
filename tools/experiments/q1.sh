export SMK_PASS_KERNEL=tma
python - <<'PY'
import sys; sys.path.insert(0,'.')
import importlib; smk = importlib.import_module('smoke-simulation_b200.binding')
print("selfcheck (bad, ties):", smk.selfcheck_omega())
print("selfcheck around 1.0:", smk.selfcheck_omega(0x3f800000, 1<<20))
PY
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -5
