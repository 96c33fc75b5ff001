set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_test25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test25.log; tail -3 gpurun_out/r2_test25.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref_n1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_ref_n1.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_bench25.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench25.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
