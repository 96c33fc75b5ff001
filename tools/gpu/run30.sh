set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/large_sweep.py --chi 512 --model hubbard --sites 16 --shard-mps --profile > gpurun_out/r2_large_small2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_large_small2.log | cut -c1-1600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/large_sweep.py --chi 1024 --model j1j2 --sites 28 --profile > gpurun_out/r2_large_small3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_large_small3.log | cut -c1-1600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/dist_check.py --chi 512 --dmrg-chi 256 > gpurun_out/r2_dist2_final.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_dist2_final.log | cut -c1-900
