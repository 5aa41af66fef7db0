# The GPU-box recipe behind profiles/r2_*: every capture only after the same command ran clean without ncu; .ncu-rep files are
# exported to CSV on the box (gpurun returns at most 64 MiB).  Usage: gpurun --timeout 2400 -- 'bash tools/capture_profiles.sh'
set -x
cap() { # name regex cmd...
  name=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o /tmp/$name "$@" > gpurun_out/r2at_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r2at_raw_$name.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null
  python tools/ncu_regions.py /tmp/$name.src.csv > gpurun_out/r2at_regions_$name.txt 2>&1
}
python tools/prof_one.py x2 65536 > gpurun_out/r2at_plain_x2.log 2>&1 && cap x2 ms_x2 python tools/prof_one.py x2 65536
python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2at_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2at_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2at_ncu_bench.log 2>&1
python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 > gpurun_out/r2at_plain_tile.log 2>&1 && cap tile ms_tile python tools/time_code.py decodeMinSum dvbs2 10 4736 f32 3.0 1
python tools/time_code.py decodeOffsetMinSum dvbs2 10 18944 f16x2 3.0 > gpurun_out/r2at_plain_tileh.log 2>&1 && cap tileh ms_tileh python tools/time_code.py decodeOffsetMinSum dvbs2 10 9472 f16x2 3.0 1
python tools/time_code.py decodeMinSum PEG 50 262144 f32 2.0 > gpurun_out/r2at_plain_quad.log 2>&1 && cap quad ms_quad python tools/time_code.py decodeMinSum PEG 50 65536 f32 2.0 1
python tools/time_code.py decodeBP 802_3_H 10 131072 f32 4.0 > gpurun_out/r2at_plain_bp.log 2>&1 && cap bp ms_rc_kernel python tools/time_code.py decodeBP 802_3_H 10 65536 f32 4.0 1
( time python bench.py > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err ) 2> gpurun_out/r2at_bench.time
du -sh gpurun_out
