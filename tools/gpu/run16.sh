set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 900 python tests/perf/sweep_bench.py c3full --split-timing > gpurun_out/r2_c3full_prof2.log 2>&1; tail -1 gpurun_out/r2_c3full_prof2.log | cut -c1-2500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py --chi 2048 --dmrg-chi 256 > gpurun_out/r2_dist2_2048b.log 2>&1; tail -1 gpurun_out/r2_dist2_2048b.log | cut -c1-2500
