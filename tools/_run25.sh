set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "f32 or BP or bp" 2>&1 | tail -5 > gpurun_out/r2y_pytest.log
python tools/time_code.py decodeBP 802_3_H 10 131072 f32 4.0 > gpurun_out/r2y_time.log 2>&1
LDPC_GPU_BP_1CTA=1 python tools/time_code.py decodeBP 802_3_H 10 131072 f32 4.0 >> gpurun_out/r2y_time.log 2>&1
LDPC_GPU_NO_RC=1 python tools/time_code.py decodeBP 802_3_H 10 131072 f32 4.0 >> gpurun_out/r2y_time.log 2>&1
