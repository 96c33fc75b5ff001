set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/large_sweep.py --chi 512 --model j1j2 --sites 24 --profile > gpurun_out/r2_large_small.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2_large_small.log | cut -c1-2000
