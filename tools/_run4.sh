set -x
python -m pytest tests/test_gpu_x2.py tests/test_gpu_lattice.py -x -q -s 2>&1 | tail -40 > gpurun_out/r2d_pytest_x2.log
python tools/time_x2.py > gpurun_out/r2d_time_x2.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_pytest_all.log
