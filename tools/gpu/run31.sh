set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n8.log | cut -c1-300
