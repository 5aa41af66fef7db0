timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_family.py -x -q -k "small_code or (min_sum_kernels_agree and PEG)" 2>&1 | tail -4 > gpurun_out/r2aw_pytest.log
python tools/time_code.py decodeMinSum PEG 50 262144 f32 2.0 > gpurun_out/r2aw_time.log 2>&1
python tools/time_code.py decodeMinSum PEG 50 524288 f32 2.0 >> gpurun_out/r2aw_time.log 2>&1
