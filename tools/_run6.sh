set -x
python -m pytest tests/test_gpu_x2.py -x -q -s 2>&1 | tail -12 > gpurun_out/r2f_pytest_x2.log
python tools/time_x2.py > gpurun_out/r2f_time_x2.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
