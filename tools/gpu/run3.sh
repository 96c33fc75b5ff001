set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q --deselect tests/test_gpu_dmrg.py::test_baseline_config2_first_sweep_against_reference > gpurun_out/r2_test3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test3.log
tail -12 gpurun_out/r2_test3.log | cut -c1-400
for i in 1 2; do timeout 600 python tools/determinism_check.py --big > gpurun_out/r2_det_big_$i.log 2>&1; done
diff gpurun_out/r2_det_big_1.log gpurun_out/r2_det_big_2.log | cut -c1-300
cut -c1-200 gpurun_out/r2_det_big_1.log
