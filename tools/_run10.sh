set -x
for m in 0 1; do LDPC_GPU_FDIV_MODE=$m python tools/prof_one.py f32 524288 10 0 nms; done > gpurun_out/r2j_fdiv.log 2>&1
LDPC_GPU_NO_FDIV=1 python tools/prof_one.py f32 524288 10 0 nms >> gpurun_out/r2j_fdiv.log 2>&1
python tools/prof_one.py f32 524288 10 0 oms >> gpurun_out/r2j_fdiv.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lattice.py -x -q 2>&1 | tail -5 >> gpurun_out/r2j_fdiv.log
