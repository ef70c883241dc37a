q() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-verify 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['ms_per_step'],3), round(r['launch_ms'],4))"; }
for g in 64 80 96 128 160 192 224; do q $g tma; SMK_PASS_KERNEL=reg q $g reg; done
