set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "small_code_multi" 2>&1 | tail -8 > gpurun_out/r2u_pytest.log
python tools/time_code.py decodeMinSum PEG 50 262144 f32 2.0 >> gpurun_out/r2u_time.log 2>&1
python tools/time_code.py decodeMinSum PEG 50 524288 f32 2.0 >> gpurun_out/r2u_time.log 2>&1
