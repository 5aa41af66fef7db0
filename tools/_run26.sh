set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2z_pytest_all.log
python tools/time_code.py decodeBP 802_3_H 10 131072 f32 4.0 > gpurun_out/r2z_time.log 2>&1
