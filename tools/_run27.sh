timeout 1500 python -m pytest tests/test_gpu_kernel_family.py tests/test_gpu_lattice.py -x -q 2>&1 | tail -15 > gpurun_out/r2aa_pytest.log
