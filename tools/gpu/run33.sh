set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tests/perf/sweep_bench.py c2 --profile > gpurun_out/r2_c2_prof2.log 2>&1; tail -1 gpurun_out/r2_c2_prof2.log | cut -c1-1300
timeout 900 python -m pytest tests/test_gpu_dmrg.py -x -q > gpurun_out/r2_test33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test33.log; tail -3 gpurun_out/r2_test33.log
timeout 600 python tests/perf/sweep_bench.py c3full --split-timing > gpurun_out/r2_c3full_prof3.log 2>&1; tail -1 gpurun_out/r2_c3full_prof3.log | cut -c1-1500
