#!/bin/bash
# A/B of the TMA pass kernel: chunk counts, against reg
q() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-verify 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', round(d['ms_per_step'],3), round(r['launch_ms'],4), round(r['frac_compulsory'],3))"; }
SMK_PASS_KERNEL=reg q C2 reg
for n in 0 5 6 7 8 10; do SMK_PASS_KERNEL=tma SMK_PASS_NCHUNKS=$n q C2 tma_n$n; done
SMK_PASS_KERNEL=reg q C3 reg
for n in 0 4 5 6 8; do SMK_PASS_KERNEL=tma SMK_PASS_NCHUNKS=$n q C3 tma_n$n; done
SMK_PASS_KERNEL=tma SMK_PASS_DEBUG=1 timeout 60 python tools/cta_times.py C2 30 2>&1 | grep -E "lean|general|busy"
SMK_PASS_KERNEL=tma SMK_PASS_NCHUNKS=8 SMK_PASS_DEBUG=1 timeout 60 python tools/cta_times.py C2 30 2>&1 | grep -E "lean|general|busy"
