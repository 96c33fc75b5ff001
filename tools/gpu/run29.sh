set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
for flag in "" "--shard-mps"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/large_sweep.py --chi 512 --model hubbard --sites 16 $flag > gpurun_out/r2_large_small$flag.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_large_small$flag.log | cut -c1-700
done
