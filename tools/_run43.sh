timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2aq_pytest_all.log
timeout 600 python bench.py --no-cpu --no-extras --workload ms_dvbs2 --steps 3 > gpurun_out/r2aq_bench_dvbs2.json 2> gpurun_out/r2aq_bench_dvbs2.err
