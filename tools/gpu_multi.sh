#!/bin/bash
# Multi-GPU check on an N-GPU box:  bash tools/gpu_multi.sh <tag> <N> [full]
# real multi-process bit-identity (tests/test_multi_gpu.py, tools/mp_check.py log), then bench.py at N GPUs
# (default workload C5 weak scaling, with its parity hashes and halo cost), and -- "full" -- C3 strong scaling and C4.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; T=${1:-mg}; N=${2:-2}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655 --nproc-per-node"
[ -z "$SKIP_PYTEST" ] && { timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/${T}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/${T}_pytest_multi.log; }
for dims in "64 48 40" "256 256 80"; do
  timeout 300 $TR ${MPW:-2} tools/mp_check.py p2p $dims 2>&1 | grep -E "mp_check|MISMATCH" | tee -a $O/${T}_mp_check.log
done
# adaptive advection margin: a hitch tick (dt = 0.5, backtraces of ~3 planes) must stay bit-identical; dt = 16 (reach beyond the
# ghost planes) must be reported as SMK_ERR_REACH by the slab ranks
timeout 300 $TR 2 tools/mp_check.py p2p 96 64 48 hitch 2>&1 | grep -E "mp_check|MISMATCH" | tee -a $O/${T}_mp_check.log
timeout 300 $TR 2 tools/mp_check.py p2p 96 64 48 reach 2>&1 | grep -E "mp_check|MISMATCH" | tee -a $O/${T}_mp_check.log
ns="$N"; [ "$3" = "full" ] && ns="1 2 4 8"
for n in $ns; do
  [ $n -gt $N ] && continue
  if [ $n -eq 1 ]; then timeout 900 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_c5_n1.json 2> $O/${T}_c5_n1.err
  else timeout 900 $TR $n bench.py --gpus $n --steps 10 --warmup 3 > $O/${T}_c5_n$n.json 2> $O/${T}_c5_n$n.err; fi
  echo "C5 n=$n rc=$?"
done
if [ "$3" = "full" ]; then
  for n in 1 2 4 8; do
    [ $n -gt $N ] && continue
    if [ $n -eq 1 ]; then timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_c3_n1.json 2> $O/${T}_c3_n1.err
    else timeout 900 $TR $n bench.py --gpus $n --workload C3 --steps 10 --warmup 3 > $O/${T}_c3_n$n.json 2> $O/${T}_c3_n$n.err; fi
    echo "C3 strong n=$n rc=$?"
  done
  timeout 1200 $TR 8 bench.py --gpus 8 --workload C4 --steps 6 --warmup 3 > $O/${T}_c4_n8.json 2> $O/${T}_c4_n8.err; echo "C4 n=8 rc=$?"
  timeout 1200 python bench.py --workload C4 --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_c4_n1.json 2> $O/${T}_c4_n1.err; echo "C4 n=1 rc=$?"
fi
for f in $O/${T}_c*_n*.json; do
python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    h=d.get("halo") or {}
    print(sys.argv[1].split("/")[-1], "n=%d ms/step=%.3f value=%.3e e2e=%.3e parity=%s halo_ms=%s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], (d.get("parity") or {}).get("result"), h.get("ms_per_step")))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
