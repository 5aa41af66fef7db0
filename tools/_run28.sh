set -x
( time python bench.py > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err ) 2> gpurun_out/r2ab_bench.time
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ab_ref.json 2> gpurun_out/r2ab_ref.err ) 2> gpurun_out/r2ab_ref.time
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ab_smoke.log 2>&1
