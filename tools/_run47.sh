timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2au_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2au_smoke.log 2>&1
