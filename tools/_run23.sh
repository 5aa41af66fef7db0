LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/libdev_fake.so python tools/time_code.py decodeMinSum PEG 50 262144 f32 2.0 > gpurun_out/r2w_fake.log 2>&1
