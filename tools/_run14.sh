set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2n_pytest_all.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1
