timeout 600 python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -3 > gpurun_out/r2ax_pytest.log
python tools/prof_one.py x2 524288 > gpurun_out/r2ax_time.log 2>&1
python tools/prof_one.py x2 524288 >> gpurun_out/r2ax_time.log 2>&1
python tools/time_e2e_dev.py 2>&1 | grep 524288 >> gpurun_out/r2ax_time.log
