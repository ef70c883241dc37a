timeout 800 python -m pytest tests -m gpu -q -x -k "not multi_gpu" 2>&1 | tail -3
q() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-verify 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['ms_per_step'],3), round(r['launch_ms'],4), round(r['frac_compulsory'],3))"; }
q C1 auto; q 128 auto; q 160 auto; SMK_PASS_KERNEL=reg q C2 reg; q C2 auto; SMK_PASS_KERNEL=tma q 160 tma
