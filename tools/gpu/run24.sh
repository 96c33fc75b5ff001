set -x
cd $GRAFT_REPO_ROOT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dist_check.py --chi 512 > gpurun_out/r2_dist2_512c.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_dist2_512c.log | python -c "
import sys, json
z = json.loads(sys.stdin.read()); print({k: z[k] for k in ('ok', 'host_path_rel_err', 'host_path_ragged_rel_err', 'host_note', 'fused_note', 'host_path_ms', 'host_path_serial_ms')})"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/dist_check.py --chi 512 --dtype c128 > gpurun_out/r2_dist2_512d.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_dist2_512d.log | cut -c1-200
