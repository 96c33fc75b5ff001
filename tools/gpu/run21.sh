set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py --chi 1024 --dtype c128 --dmrg-chi 256 > gpurun_out/r2_dist2_c128.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_dist2_c128.log | cut -c1-2800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/dist_check.py --chi 512 --dtype f64 --dmrg-chi 128 > gpurun_out/r2_dist2_512b.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_dist2_512b.log | cut -c1-600
