set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/ozaki_experiment.py > gpurun_out/r2_ozaki.log 2>&1; cat gpurun_out/r2_ozaki.log | cut -c1-600
