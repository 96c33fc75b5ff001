set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python tests/perf/sweep_bench.py c3multi > gpurun_out/r2_c3multi.log 2>&1; cat gpurun_out/r2_c3multi.log | cut -c1-700
