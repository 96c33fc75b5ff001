set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/first_bond_diag.py --chi 2048 --L 100 > gpurun_out/r2_fb_1.log 2>&1; tail -20 gpurun_out/r2_fb_1.log | cut -c1-1500
timeout 600 python tools/first_bond_diag.py --chi 2048 --L 100 > gpurun_out/r2_fb_2.log 2>&1
timeout 600 python tools/first_bond_diag.py --chi 2048 --L 26 > gpurun_out/r2_fb_3.log 2>&1
