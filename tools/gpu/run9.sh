set -x
cd $GRAFT_REPO_ROOT
timeout 900 python tests/perf/sweep_bench.py c3full --split-timing > gpurun_out/r2_c3full_prof.log 2>&1; tail -1 gpurun_out/r2_c3full_prof.log
