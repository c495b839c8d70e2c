# round-end GPU batch: full -m gpu suite + the default bench line (outputs under gpurun_out/r02f)
set -x
O=gpurun_out/r02f; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; tail -c 300 $O/bench.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
