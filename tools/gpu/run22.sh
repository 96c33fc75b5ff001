set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_test22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test22.log; tail -3 gpurun_out/r2_test22.log
python tools/profile_aux.py > gpurun_out/r02_aux_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lz_|mix_apply|mix_prepare|identity_dev" -c 16 -o gpurun_out/r02_prof_aux python tools/profile_aux.py > gpurun_out/r02_ncu_aux.log 2>&1
tail -2 gpurun_out/r02_ncu_aux.log
