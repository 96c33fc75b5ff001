set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 tools/large_sweep.py --chi 1024 --model hubbard --profile > gpurun_out/r2_large_c5.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2_large_c5.log | cut -c1-2500
