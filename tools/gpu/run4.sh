set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do timeout 900 python tools/determinism_check.py --big --chi 2048 --L 26 > gpurun_out/r2_det_c3_$i.log 2>&1; tail -1 gpurun_out/r2_det_c3_$i.log | cut -c1-900; done
