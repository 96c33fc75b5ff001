set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tests/perf/sweep_bench.py split > gpurun_out/r2_split.log 2>&1; cat gpurun_out/r2_split.log | cut -c1-1800
timeout 900 python tests/perf/sweep_bench.py c3full > gpurun_out/r2_c3full_sk.log 2>&1; tail -1 gpurun_out/r2_c3full_sk.log
