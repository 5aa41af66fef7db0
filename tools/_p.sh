( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2bf_pytest_all.log 2>&1
python tools/time_code.py decodeOffsetMinSum dvbs2 10 18944 f16x2 3.0 > gpurun_out/r2bf_time.log 2>&1
python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 >> gpurun_out/r2bf_time.log 2>&1
python tools/time_code.py decodeMinSum dvbs2 10 2368 f64 3.0 >> gpurun_out/r2bf_time.log 2>&1
python tools/prof_one.py x2 524288 >> gpurun_out/r2bf_time.log 2>&1
