set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_lanczos.py -x -q > gpurun_out/r2_test23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test23.log; tail -2 gpurun_out/r2_test23.log
python tools/profile_aux.py > gpurun_out/r02_aux_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lz_" -c 12 -o gpurun_out/r02_prof_aux2 python tools/profile_aux.py > gpurun_out/r02_ncu_aux2.log 2>&1
tail -2 gpurun_out/r02_ncu_aux2.log
