#!/usr/bin/env python3
"""Is torch symmetric memory usable beyond 2 GiB?  Each rank writes a pattern into its peer's buffer at several
offsets (peer-mapped pointer, cudaMemcpy through torch) and the owner verifies it.  torchrun --nproc-per-node 2."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world, local = bench.dist_setup(0)
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
total = int(gib * (1 << 30))
t = symm_mem.empty(total, dtype=torch.uint8, device=dev)
t.zero_()
torch.cuda.synchronize()
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
peer = (rank + 1) % world
res = {}
for off_gib in (0.25, 1.0, 1.9, 2.1, gib - 0.1):
    off = int(off_gib * (1 << 30)) // 256 * 256
    n = 1 << 20
    try:
        buf = hdl.get_buffer(peer, (n,), torch.uint8, off)          # storage_offset in elements of uint8
        buf.fill_(17 + rank)
        torch.cuda.synchronize()
        dist.barrier()
        mine = t[off: off + n]
        ok = bool((mine == 17 + (rank - 1) % world).all())
        res[off_gib] = ok
    except Exception as ex:
        res[off_gib] = repr(ex)[:100]
    dist.barrier()
print(f"rank {rank}: total {gib} GiB ->", res, "ptrs", [hex(int(p)) for p in hdl.buffer_ptrs], flush=True)
dist.destroy_process_group()
