timeout 300 python -m pytest tests -x -q -m gpu -k "dvbs2 or tile or hbm or HBM" 2>&1 | tail -6 > gpurun_out/r2ao_pytest.log
timeout 120 python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 > gpurun_out/r2ao_time.log 2>&1
timeout 120 python tools/time_code.py decodeMinSum dvbs2 10 9472 f64 3.0 >> gpurun_out/r2ao_time.log 2>&1
LDPC_GPU_NO_TILEC=1 timeout 120 python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 >> gpurun_out/r2ao_time.log 2>&1
