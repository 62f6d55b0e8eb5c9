mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short > gpurun_out/r2aq_pytest_all.log 2>&1
tail -3 gpurun_out/r2aq_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
