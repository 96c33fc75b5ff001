set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py --chi 512 > gpurun_out/r2_dist2_512.log 2>&1; tail -2 gpurun_out/r2_dist2_512.log | cut -c1-3000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py --chi 2048 --dmrg-chi 256 > gpurun_out/r2_dist2_2048.log 2>&1; tail -2 gpurun_out/r2_dist2_2048.log | cut -c1-3000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep > gpurun_out/r2_bench_n2.log 2>&1; tail -1 gpurun_out/r2_bench_n2.log | cut -c1-2500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --serial-e2e --no-extras > gpurun_out/r2_bench_n2_serial.log 2>&1; tail -1 gpurun_out/r2_bench_n2_serial.log | cut -c1-1500
