export SMK_PASS_KERNEL=tma SMK_PASS_DEBUG=1 TRACE_ALL=1
SMK_TRACE_TILE=2,5,2 timeout 60 python tools/cta_times.py C2 20 2>&1 | tail -48
python -m pytest tests/test_parity_gaps_gpu.py -q -x -k omega 2>&1 | tail -30
