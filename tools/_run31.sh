python bench.py --no-cpu --no-extras --e2e-frames 1048576 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err
python tools/time_e2e_dev.py > gpurun_out/r2ae_e2e.log 2>&1
