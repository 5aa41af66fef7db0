for lib in libldpc_gpu libdev_st1800 libdev_st20000 libldpc_gpu; do
  LDPC_GPU_LIB=$PWD/ldpcsimulation_b200/_build/$lib.so python tools/prof_one.py x2 524288 >> gpurun_out/r2af_stagger.log 2>&1
done
