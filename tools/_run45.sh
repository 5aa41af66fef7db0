for lib in libldpc_gpu libdev_nosplit; do
  export LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/$lib.so
  echo "== $lib" >> gpurun_out/r2as_cmp.log
  python tools/prof_one.py x2 524288 2>&1 | cut -c1-50 >> gpurun_out/r2as_cmp.log
  python tools/time_code.py decodeMinSum PEG 50 262144 f32 2.0 2>&1 | cut -c1-120 >> gpurun_out/r2as_cmp.log
  python tools/time_code.py decodeBP 802_3_H 10 131072 f32 4.0 2>&1 | cut -c1-120 >> gpurun_out/r2as_cmp.log
  python tools/time_code.py decodeNormalizedMinSum 802_3_H 10 262144 f32 4.0 2>&1 | cut -c1-120 >> gpurun_out/r2as_cmp.log
  python tools/time_code.py decodeMinSum dvbs2 10 9472 f32 3.0 2>&1 | cut -c1-120 >> gpurun_out/r2as_cmp.log
  python tools/time_e2e_dev.py 2>&1 | grep "524288" >> gpurun_out/r2as_cmp.log
done
