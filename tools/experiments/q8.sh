D=smoke-simulation_b200
q() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-verify 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['ms_per_step'],3), round(r['launch_ms'],4), round(r['frac_compulsory'],3))"; }
for v in a b a b; do cp $D/variant_$v.so.bin $D/libsmoke_b200.so; q C2 $v; q C3 $v; 
SMK_PASS_DEBUG=1 timeout 60 python tools/cta_times.py C2 20 2>&1 | grep -E "lean|general"; done
cp $D/variant_b.so.bin $D/libsmoke_b200.so; timeout 300 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -2
