set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do timeout 1200 python tools/determinism_check.py --sweep-only --chi 2048 --L 100 > gpurun_out/r2_det_c3full_$i.log 2>&1; tail -1 gpurun_out/r2_det_c3full_$i.log | cut -c1-1200; done
