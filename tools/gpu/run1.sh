set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dmrg.py::test_baseline_config2_first_sweep_against_reference > gpurun_out/r2_test1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test1.log
tail -15 gpurun_out/r2_test1.log
timeout 900 python tools/determinism_check.py --chi 512 --L 16 > gpurun_out/r2_determinism.log 2>&1; echo "rc=$?" >> gpurun_out/r2_determinism.log
cat gpurun_out/r2_determinism.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-sweep > gpurun_out/r2_bench_canon.log 2>&1; tail -c 3000 gpurun_out/r2_bench_canon.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --dense-env > gpurun_out/r2_bench_dense.log 2>&1; tail -c 1500 gpurun_out/r2_bench_dense.log
