set -x
timeout 600 python -m pytest tests/test_sc_ngdbf.py tests/test_gpu_replay.py -x -q 2>&1 | tail -15 > gpurun_out/r2m_pytest_sc.log
