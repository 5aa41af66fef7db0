( time python bench.py --no-cpu > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err ) 2> gpurun_out/r2ad_bench.time
python bench.py --no-cpu --no-extras --workload ngdbfhw_8023 --steps 3 > gpurun_out/r2ad_bench_hw.json 2> gpurun_out/r2ad_bench_hw.err
python bench.py --no-cpu --no-extras --workload ms_peg_t50 --steps 3 > gpurun_out/r2ad_bench_peg.json 2> gpurun_out/r2ad_bench_peg.err
