set -x
export LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/libdev_fake.so
python tools/prof_one.py x2 262144 > gpurun_out/r2t_fake.log 2>&1
LDPC_GPU_X2_VARIANT=2 python tools/prof_one.py x2 262144 >> gpurun_out/r2t_fake.log 2>&1
unset LDPC_GPU_LIB
python tools/prof_one.py x2 262144 >> gpurun_out/r2t_fake.log 2>&1
LDPC_GPU_X2_VARIANT=2 python tools/prof_one.py x2 262144 >> gpurun_out/r2t_fake.log 2>&1
