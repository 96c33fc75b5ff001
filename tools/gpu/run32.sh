set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tests/perf/sweep_bench.py c2 > gpurun_out/r2_c2.log 2>&1; tail -1 gpurun_out/r2_c2.log | cut -c1-1500
timeout 600 python tests/perf/sweep_bench.py c2 --profile > gpurun_out/r2_c2_prof.log 2>&1; tail -1 gpurun_out/r2_c2_prof.log | cut -c1-900
