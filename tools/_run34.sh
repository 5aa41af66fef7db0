timeout 900 python -m pytest tests/test_gpu_x2.py tests/test_gpu_errors.py -x -q 2>&1 | tail -4 > gpurun_out/r2ah_pytest.log
python bench.py --no-cpu --no-extras > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err
