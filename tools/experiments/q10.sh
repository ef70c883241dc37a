# 8-GPU box: C5 weak at N = 8 and C3 strong at N = 8, default kernels, with the bit-identity check against one GPU
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655 --nproc-per-node"
timeout 600 $TR 8 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2b_scale_c5_n8.json 2> $O/r2b_scale_c5_n8.err; echo "C5 n=8 rc=$?"
timeout 300 $TR 8 bench.py --gpus 8 --workload C3 --steps 10 --warmup 3 > $O/r2b_scale_c3_n8.json 2> $O/r2b_scale_c3_n8.err; echo "C3 n=8 rc=$?"
for f in $O/r2b_scale_c5_n8.json $O/r2b_scale_c3_n8.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); h=d.get("halo") or {}
    print(sys.argv[1].split("/")[-1], "n=%d ms/step=%.3f value=%.3e e2e=%.3e parity=%s halo_ms=%s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], (d.get("parity") or {}).get("result"), h.get("ms_per_step")))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
