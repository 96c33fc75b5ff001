set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tests/perf/sweep_bench.py c2 --profile > gpurun_out/r2_c2_prof3.log 2>&1; tail -1 gpurun_out/r2_c2_prof3.log | cut -c1-800
timeout 600 python -m pytest tests/test_gpu_dmrg.py -x -q -k "config2 or c3s128 or model_families or per_update" > gpurun_out/r2_test34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test34.log; tail -3 gpurun_out/r2_test34.log
