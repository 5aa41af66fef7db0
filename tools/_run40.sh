LDPC_GPU_FORCE_HBM_STATE=1 python tools/time_code.py decodeMinSum PEG 10 37888 f32 2.0 > gpurun_out/r2an_l2.log 2>&1
LDPC_GPU_FORCE_HBM_STATE=1 python tools/time_code.py decodeMinSum PEG 50 37888 f32 2.0 >> gpurun_out/r2an_l2.log 2>&1
LDPC_GPU_FORCE_HBM_STATE=1 python tools/time_code.py decodeMinSum 802_3_H 10 37888 f32 4.0 >> gpurun_out/r2an_l2.log 2>&1
LDPC_GPU_FORCE_HBM_STATE=1 python tools/time_code.py decodeMinSum 4000 10 37888 f32 2.8 >> gpurun_out/r2an_l2.log 2>&1
