set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
python tools/time_iters.py 262144 f32 > gpurun_out/r2a_time_f32.log 2>&1
LDPC_GPU_NO_FDIV=1 python tools/time_iters.py 262144 f32 > gpurun_out/r2a_time_f32_nofdiv.log 2>&1
python tools/time_iters.py 262144 f16x2 > gpurun_out/r2a_time_h2.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_smi.log
lscpu | head -25 > gpurun_out/r2a_lscpu.log; numactl -H >> gpurun_out/r2a_lscpu.log 2>&1; nproc >> gpurun_out/r2a_lscpu.log
