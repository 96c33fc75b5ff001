set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 6 -c 4 -o gpurun_out/r02_prof_gemm python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_ncu_f.log 2>&1
ls -la gpurun_out/ | tail -5
tail -3 gpurun_out/r02_ncu_f.log
