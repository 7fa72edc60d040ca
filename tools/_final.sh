python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r01o_bench_cfg4.json 2> gpurun_out/r01o_bench_cfg4.err; tail -c 300 gpurun_out/r01o_bench_cfg4.json; wc -l gpurun_out/r01o_bench_cfg4.json
QOC_PURE_STATE=0 python tools/time_big_ensemble.py 5:64:500 6:32:500
