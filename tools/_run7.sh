set -x
python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -5 > gpurun_out/r2g_pytest_x2.log
python tools/time_e2e.py > gpurun_out/r2g_time_e2e.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2g_pytest_all.log
