timeout 900 python -m pytest tests -x -q -m gpu -k "dvbs2 or tile or hbm or HBM" 2>&1 | tail -4 > gpurun_out/r2am_pytest.log
python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 > gpurun_out/r2am_time.log 2>&1
python tools/time_code.py decodeMinSum dvbs2 10 9472 f64 3.0 >> gpurun_out/r2am_time.log 2>&1
export LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/libdev_t128.so
timeout 900 python -m pytest tests -x -q -m gpu -k "dvbs2 or tile or hbm or HBM" 2>&1 | tail -4 >> gpurun_out/r2am_pytest.log
python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 >> gpurun_out/r2am_time.log 2>&1
python tools/time_code.py decodeMinSum dvbs2 10 9472 f64 3.0 >> gpurun_out/r2am_time.log 2>&1
