export SMK_PASS_KERNEL=tma SMK_PASS_DEBUG=1
for n in 2 5 10 20; do echo "ticks $n"; timeout 60 python tools/cta_times.py C2 $n 2>&1 | grep -E "lean|general"; done
