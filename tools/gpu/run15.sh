set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
N=${NG:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n$N.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n$N.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench_ref_n$N.log | cut -c1-600
