set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n8.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-extras > gpurun_out/r2_bench_n8_nccl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n8_nccl.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/sharded_sweep.py --L 32 --chi 2048 --profile > gpurun_out/r2_sharded_sweep8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_sharded_sweep8.log | cut -c1-1500
