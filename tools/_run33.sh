python tools/time_code.py decodeMinSum 4376 10 65536 f32 5.0 > gpurun_out/r2ag_codes.log 2>&1
python tools/time_code.py decodeMinSum 4000 10 65536 f32 2.8 >> gpurun_out/r2ag_codes.log 2>&1
python tools/time_code.py decodeMinSum 4000 50 65536 f32 2.8 >> gpurun_out/r2ag_codes.log 2>&1
python tools/time_code.py decodeOffsetMinSum 802_3 10 262144 f32 4.0 >> gpurun_out/r2ag_codes.log 2>&1
python tools/time_code.py decodeNormalizedMinSum 802_3_H 10 262144 f32 4.0 >> gpurun_out/r2ag_codes.log 2>&1
