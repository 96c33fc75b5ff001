set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q --deselect tests/test_gpu_dmrg.py::test_baseline_config2_first_sweep_against_reference > gpurun_out/r2_test2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test2.log
tail -40 gpurun_out/r2_test2.log | cut -c1-600
timeout 900 python tools/determinism_check.py --chi 512 --L 16 > gpurun_out/r2_determinism_b.log 2>&1
for i in 1 2; do timeout 600 python tests/perf/sweep_bench.py c3full > gpurun_out/r2_c3full_$i.log 2>&1; tail -2 gpurun_out/r2_c3full_$i.log; done
