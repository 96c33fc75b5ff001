set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-sweep > gpurun_out/r2_bench_n2b.log 2>&1; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/dist_check.py --chi 2048 --dmrg-chi 256 > gpurun_out/r2_dist2_2048c.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_dist2_2048c.log | cut -c1-300
