set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/large_sweep.py --chi 4096 --model j1j2 --profile > gpurun_out/r2_large_c4.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2_large_c4.log | cut -c1-2500
