timeout 900 python -m pytest tests/test_gpu_x2.py tests/test_gpu_parity.py -x -q -k "x2 or quantiser or packed or QP or level" 2>&1 | tail -4 > gpurun_out/r2ai_pytest.log
python bench.py --no-cpu --no-extras > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err
python tools/time_e2e_dev.py > gpurun_out/r2ai_e2e.log 2>&1
