set -x
python tools/prof_one.py x2 65536 > gpurun_out/r2c_plain_x2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ms_x2 -s 2 -c 1 -f -o gpurun_out/r2c_x2 python tools/prof_one.py x2 65536 > gpurun_out/r2c_ncu_x2.log 2>&1
python tools/prof_one.py h2rc 65536 > gpurun_out/r2c_plain_h2rc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ms_h2rc -s 2 -c 1 -f -o gpurun_out/r2c_h2rc python tools/prof_one.py h2rc 65536 > gpurun_out/r2c_ncu_h2rc.log 2>&1
ls -la gpurun_out/*.ncu-rep
