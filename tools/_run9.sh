set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lattice.py tests/test_gpu_random_codes.py -x -q -k "802_3 or other_codes or random" 2>&1 | tail -8 > gpurun_out/r2i_pytest.log
python tools/time_code.py decodeNormalizedMinSum 802_3 10 524288 f32 4.0 > gpurun_out/r2i_time_8023.log 2>&1
python tools/time_code.py decodeNormalizedMinSum 802_3 10 131072 f64 4.0 >> gpurun_out/r2i_time_8023.log 2>&1
LDPC_GPU_NO_RC=1 python tools/time_code.py decodeNormalizedMinSum 802_3 10 524288 f32 4.0 >> gpurun_out/r2i_time_8023.log 2>&1
