set -x
python tools/prof_one.py x2 65536 > gpurun_out/r2ar_plain_x2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ms_x2 -s 2 -c 1 -f -o gpurun_out/r2ar_x2 python tools/prof_one.py x2 65536 > gpurun_out/r2ar_ncu_x2.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2ar_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ar_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2ar_ncu_bench.log 2>&1
python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 > gpurun_out/r2ar_plain_tile.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ms_tile -s 1 -c 1 -f -o gpurun_out/r2ar_tile python tools/time_code.py decodeMinSum dvbs2 10 4736 f32 3.0 1 > gpurun_out/r2ar_ncu_tile.log 2>&1
