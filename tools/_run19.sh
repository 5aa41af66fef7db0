set -x
timeout 600 python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -4 > gpurun_out/r2s_pytest.log
python tools/prof_one.py x2 524288 > gpurun_out/r2s_time.log 2>&1
python tools/prof_one.py x2 524288 >> gpurun_out/r2s_time.log 2>&1
python tools/prof_one.py x2 1048576 >> gpurun_out/r2s_time.log 2>&1
