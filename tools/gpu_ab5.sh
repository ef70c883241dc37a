#!/bin/bash
# parity suite with the TMA pass + timing A/B against reg
export SMK_PASS_KERNEL=tma
timeout 400 python -m pytest tests -m gpu -q -x -k "not multi_gpu" 2>&1 | tail -3
q() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-verify 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['ms_per_step'],3), round(r['launch_ms'],4), round(r['frac_compulsory'],3))"; }
q C2 tma; SMK_PASS_NCHUNKS=8 q C2 tma8; q C3 tma; SMK_PASS_NCHUNKS=5 q C3 tma5; SMK_PASS_KERNEL=reg q C2 reg
SMK_PASS_DEBUG=1 timeout 60 python tools/cta_times.py C2 30 2>&1 | grep -E "lean|general|busy"
