set -x
timeout 600 python -m pytest tests/test_gpu_replay.py -x -q 2>&1 | tail -25 > gpurun_out/r2h_pytest_replay.log
