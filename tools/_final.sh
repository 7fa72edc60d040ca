set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r01m_bench_cfg4.json 2> gpurun_out/r01m_bench_cfg4.err; tail -c 600 gpurun_out/r01m_bench_cfg4.json
python bench.py --config cfg5 --steps 5 > gpurun_out/r01m_bench_cfg5.json 2> gpurun_out/r01m_bench_cfg5.err; tail -c 300 gpurun_out/r01m_bench_cfg5.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -c 400
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01m_launches_cfg5.csv python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01m_launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/time_pure.py 8 256 2>/dev/null | head -1
