set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_test11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test11.log; tail -16 gpurun_out/r2_test11.log
timeout 900 python bench.py > gpurun_out/r2_bench11.log 2>&1; tail -1 gpurun_out/r2_bench11.log | cut -c1-600
