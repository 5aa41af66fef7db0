set -x
python tools/prof_one.py x2 65536 > gpurun_out/r2e_plain_x2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ms_x2 -s 2 -c 1 -f -o gpurun_out/r2e_x2 python tools/prof_one.py x2 65536 > gpurun_out/r2e_ncu_x2.log 2>&1
