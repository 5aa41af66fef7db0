set -x
ncu --set full --clock-control none --import-source on -k regex:ms_quad -s 1 -c 1 -f -o gpurun_out/r2v_quad python tools/time_code.py decodeMinSum PEG 50 65536 f32 2.0 1 > gpurun_out/r2v_ncu.log 2>&1
