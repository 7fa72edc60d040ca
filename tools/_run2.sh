python -m pytest tests/test_gpu_big.py tests/test_gpu_reference_scenarios.py -x -q 2>&1 | tail -3
export QOC_PURE_STATE=0
python tools/time_config5.py 5 6 7 8
echo general; QOC_BIG_HERM=0 python tools/time_config5.py 5 6 7 8
python tools/time_big_ensemble.py 5:64:500 6:32:500 7:8:500
