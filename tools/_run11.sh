set -x
timeout 900 python -m pytest tests/test_gpu_x2.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/r2k_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
