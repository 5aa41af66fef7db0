set -x
timeout 600 python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -4 > gpurun_out/r2p_pytest.log
python tools/prof_one.py x2 524288 > gpurun_out/r2p_time.log 2>&1
python tools/prof_one.py x2 524288 10 1 >> gpurun_out/r2p_time.log 2>&1
LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/libldpc_gpu_hifma.so python tools/prof_one.py x2 524288 >> gpurun_out/r2p_time.log 2>&1
LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/libldpc_gpu_hifma.so timeout 600 python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -3 >> gpurun_out/r2p_pytest.log
