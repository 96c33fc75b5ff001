set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_test7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test7.log; tail -5 gpurun_out/r2_test7.log
for i in 1 2; do timeout 600 python tools/determinism_check.py --sweep-only --chi 2048 --L 100 --seed-global 1 > gpurun_out/r2_det_seeded_$i.log 2>&1; tail -1 gpurun_out/r2_det_seeded_$i.log | cut -c1-900; done
