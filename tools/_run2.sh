python -m pytest tests/test_gpu_big.py -x -q 2>&1 | tail -3
python tools/time_pure.py 8 2>/dev/null | head -1; python tools/time_pure.py 6 1 2>/dev/null | head -1; python tools/time_pure.py 5 1 2>/dev/null | head -1
export QOC_PURE_STATE=0
python tools/time_config5.py 5 6 7 8
python tools/time_big_ensemble.py 5:64:500 6:32:500 7:8:500
