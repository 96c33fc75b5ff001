set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tests/perf/sweep_bench.py split > gpurun_out/r2_split2.log 2>&1; cat gpurun_out/r2_split2.log | cut -c1-2600
