"""Multi-GPU sharding of the DMRG inner loop over the left bond index.

One process per GPU (torchrun), ``torch.distributed`` over NCCL/NVLink for the plumbing.  The matvec partitions
naturally over the output/left bond index a' (rows of ``L[w, a', a]``): rank r owns the contiguous row block
``[lo_r, hi_r)`` of L, of H_eff.theta and of every Lanczos vector -- theta is row-major with chi_l outermost
(reference dmrg.py:217), so a row block is one contiguous chunk of n/P elements.  ``R`` and the MPO tensors are
replicated.  Per matvec there is exactly ONE exchange step, the all-gather of the input vector; GEMM -> mix -> GEMM
then run on the local rows with no further communication.  Lanczos needs one small all-reduce per scalar
(alpha, beta^2, PRO overlaps), done by ``eigensolver._lanczos_impl`` on the device scalar array.

The reference has no counterpart (single process, single device).
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

from .eigensolver import _lanczos_impl, _CUDA_OPS

__all__ = ["row_partition", "fused_eligible", "ShardedApplyMPO", "lanczos_ground_state_sharded", "shard_rows", "gather_rows",
           "ShardedDMRG", "CudaEngine", "FusedShardedApplyMPO", "SymmetricArena"]


def row_partition(chi: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row blocks [lo, hi) of the left bond, as even as possible (first chi % world ranks get one more)."""
    if world < 1 or chi < 0:
        raise ValueError("row_partition: need world >= 1 and chi >= 0")
    base, extra = divmod(chi, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_rows(t: torch.Tensor, dim: int, rank: int, world: int) -> torch.Tensor:
    """This rank's block of ``t`` along ``dim`` (a view)."""
    lo, hi = row_partition(t.shape[dim], world)[rank]
    return t.narrow(dim, lo, hi - lo)


def gather_rows(shard: torch.Tensor, chi: int, per_row: int, group: Any = None) -> torch.Tensor:
    """All-gather flat row shards (rows x per_row elements each) into the full flat vector."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = row_partition(chi, world)
    full = torch.empty(chi * per_row, dtype=shard.dtype, device=shard.device)
    if all((hi - lo) == (parts[0][1] - parts[0][0]) for lo, hi in parts):
        dist.all_gather_into_tensor(full, shard.contiguous(), group=group)
    else:
        _uneven_gather(full, shard, parts, per_row, group)
    return full


def _uneven_gather(full: torch.Tensor, shard: torch.Tensor, parts, per_row: int, group: Any) -> None:
    """Row blocks of different sizes (chi not divisible by the world size): one broadcast per owner."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    lo, hi = parts[rank]
    full[lo * per_row: hi * per_row].copy_(shard.reshape(-1))
    for r, (a, b) in enumerate(parts):
        if b > a:
            dist.broadcast(full[a * per_row: b * per_row], src=dist.get_global_rank(group, r) if group is not None
                           else r, group=group)


class ShardedApplyMPO:
    """H_eff . theta with L row-sharded over the ranks of ``group``.

    ``local_op(psi_full, L_rows, M1, M2, R, out=...)`` computes the rows of H_eff.theta this rank owns; by default
    it is the CUDA ``ApplyMPO`` of this package.  Call signature follows the Lanczos operator protocol
    (lanczos_solver.py:234-238): ``op(v_shard, *args, out=w_shard)`` with args = (L_rows, M1, M2, R).
    """

    def __init__(self, chi_l: int, per_row: int, group: Any = None, local_op: Optional[Callable] = None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.chi_l, self.per_row = chi_l, per_row
        self.lo, self.hi = row_partition(chi_l, self.world)[self.rank]
        self._even = chi_l % self.world == 0
        self._full: Optional[torch.Tensor] = None
        if local_op is None:
            from .dmrg import ApplyMPO
            local_op = ApplyMPO()
        self.local_op = local_op

    def gather(self, v_shard: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        n = self.chi_l * self.per_row
        if self._full is None or self._full.numel() != n or self._full.dtype != v_shard.dtype \
                or self._full.device != v_shard.device:
            self._full = torch.empty(n, dtype=v_shard.dtype, device=v_shard.device)
        if self._even:
            dist.all_gather_into_tensor(self._full, v_shard, group=self.group)
        else:
            _uneven_gather(self._full, v_shard, row_partition(self.chi_l, self.world), self.per_row, self.group)
        return self._full

    def __call__(self, v_shard, L_rows, M1, M2, R, out=None):
        full = self.gather(v_shard)                      # the one exchange step of the sharded matvec
        return self.local_op(full, L_rows, M1, M2, R, out=out)


class SymmetricArena:
    """One peer-mapped (``torch.distributed._symmetric_memory``) allocation shared by every fused operator of a
    process group: [flags (256 B)] [gathered buffer 0] [gathered buffer 1].  Allocating and rendezvous-ing symmetric
    memory is a slow collective, so a sweep does it once for the largest bond instead of once per bond shape.  The
    epoch counter is shared too (flags are monotonic per source rank)."""

    def __init__(self, max_vector_bytes: int, device, group: Any = None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.stride = (int(max_vector_bytes) + 255) // 256 * 256
        total = 256 + 2 * self.stride
        try:
            symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self.store = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.store.zero_()
        torch.cuda.synchronize(device)
        self.hdl = symm_mem.rendezvous(self.store, self.group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.flags = self.store[:256].view(torch.int32)
        self.flag_ptrs = [self.ptrs[r] + 4 * self.rank for r in range(self.world)]
        # second half of the flag block: the (source rank x column block) grid of the host-resident path
        self.grid_cols = max(1, min(8, 32 // self.world))
        self.grid_flags = self.flags[32:]
        self.copy_streams = {"h2d": torch.cuda.Stream(device=device), "d2h": torch.cuda.Stream(device=device)}
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        self.side = [torch.cuda.Stream(device=device) for _ in range(2)]
        self.epoch = 0
        dist.barrier(group=self.group)


def fused_eligible(chi_l: int, per_row: int, dtype: torch.dtype, world: int) -> bool:
    """Can the arrival-gated (TMA-staged) step-1 GEMM describe this bond?  The k range (chi_l) must split into
    ``world`` chunks of whole BK-deep k-tiles, and theta viewed as (chi_l, per_row) -- the operand whose ROW index is
    contiguous -- needs per_row to be a multiple of one 128-byte line (16 float64 / 8 complex128), see
    csrc/gemm_tma.cu ``tma_eligible``.  A cutoff-truncated or spin-1 right bond (e.g. chi_r = 50 or 27) fails the
    second condition; such bonds take the all-gather + plain matvec path (``ShardedApplyMPO``)."""
    bk = 8 if dtype.is_complex else 16
    line = 8 if dtype.is_complex else 16
    return world >= 1 and chi_l % (bk * world) == 0 and per_row % line == 0


class FusedShardedApplyMPO:
    """Sharded H_eff . theta with the all-gather FUSED into the matvec (no NCCL call on the data path).

    Every rank pushes its row shard of the input vector straight into all peers' gathered buffers with P2P stores
    over NVLink (``qsb_push_shard``: one kernel, then a system-scope release of a per-source arrival flag).  The
    step-1 GEMM then walks its k range source-rank block by source-rank block, starting with the rank's own block
    (already local) and touching block c only once its flag shows it has landed -- the transfer overlaps the first
    tiles instead of preceding the kernel (``qsb_heff_desc.n_chunks / chunk_flags``).  The gathered buffers and flags
    live in ``torch.distributed._symmetric_memory`` (peer-mapped allocations); buffers are double-buffered by epoch
    parity, flags are monotonic epochs, so a rank that runs one step ahead never overwrites data a peer still reads.
    """

    def __init__(self, chi_l: int, per_row: int, dtype: torch.dtype, device, group: Any = None, push: str = "dma",
                 arena: Optional[SymmetricArena] = None):
        import torch.distributed as dist
        from . import _lib
        self._ops = _lib.ops
        self._lib = _lib
        if push not in ("dma", "sm"):
            raise ValueError("push must be 'dma' (copy engines on side streams) or 'sm' (one store kernel)")
        self.push = push
        grp = group if group is not None else dist.group.WORLD
        world = dist.get_world_size(grp)
        es = torch.empty(0, dtype=dtype).element_size()
        bk = 8 if dtype.is_complex else 16
        if not fused_eligible(chi_l, per_row, dtype, world):
            raise ValueError(f"fused sharded matvec needs chi_l divisible by {bk} * world_size and d1*d2*chi_r "
                             f"divisible by {bk} (got chi_l={chi_l}, per_row={per_row}, world={world}); "
                             "use ShardedApplyMPO")
        n = chi_l * per_row
        if arena is None:
            arena = SymmetricArena(n * es, device, grp)
        elif arena.stride < n * es:
            raise ValueError("symmetric arena too small for this bond")
        self.arena = arena
        self.group, self.world, self.rank = arena.group, arena.world, arena.rank
        self.chi_l, self.per_row = chi_l, per_row
        self.lo, self.hi = row_partition(chi_l, self.world)[self.rank]
        shard_bytes = (n // self.world) * es
        st = arena.stride
        self.flags = arena.flags
        self.bufs = [arena.store[256 + b * st: 256 + b * st + n * es].view(dtype) for b in (0, 1)]
        self.dst_ptrs = [[arena.ptrs[r] + 256 + b * st + self.rank * shard_bytes for r in range(self.world)]
                         for b in (0, 1)]
        self.flag_ptrs = arena.flag_ptrs
        self.ticket = arena.ticket
        self._side = arena.side if push == "dma" else []
        # push order: own block, then rank-1, rank-2, ...  Rank m therefore RECEIVES blocks m+1, m+2, ... in that
        # order -- the order in which its gated GEMM walks the k chunks (rotation starts at its own block).
        self._order = [(self.rank - i) % self.world for i in range(self.world)]

    @property
    def epoch(self) -> int:
        return self.arena.epoch

    def __call__(self, v_shard, L_rows, M1, M2, R, out=None):
        self.arena.epoch += 1
        b = self.arena.epoch & 1
        v = v_shard.reshape(-1)
        if not v.is_contiguous():
            v = v.contiguous()
        if out is None:
            out = torch.empty(L_rows.shape[1] * M1.shape[2] * M2.shape[2] * R.shape[1], dtype=v.dtype, device=v.device)
        (li, lrow0), (ri, _r0) = self._lib.identity_of(L_rows), self._lib.identity_of(R)   # verified identity channels
        if self.push == "sm":
            self._ops.push_shard(v, self.dst_ptrs[b], self.flag_ptrs, self.epoch, self.ticket)
            self._ops.heff_apply_gated(self.bufs[b], L_rows, M1, M2, R, out, self.flags, self.world, self.rank,
                                       self.epoch, 0, li, ri, lrow0)
            return out
        # copy engines on a side stream: the pushes run WHILE the gated GEMM already works on the local block
        main = torch.cuda.current_stream(v.device)
        # own block: a plain stream-ordered copy on the main stream (a same-device copy may run on SMs, which the
        # persistent GEMM would be holding while it waited for the flag -> never gate the local block on a flag)
        nloc = v.numel()
        self.bufs[b][self.rank * nloc:(self.rank + 1) * nloc].copy_(v)
        ready = torch.cuda.Event()
        ready.record(main)
        dones = []
        peers = self._order[1:]
        for si, side in enumerate(self._side):       # two copy streams: peer destinations dealt alternately
            targets = peers[si::len(self._side)]
            if not targets:
                continue
            side.wait_event(ready)
            with torch.cuda.stream(side):
                self._lib.push_shard_dma(v, [self.dst_ptrs[b][r] for r in targets],
                                         [self.flag_ptrs[r] for r in targets], self.epoch)
                done = torch.cuda.Event()
                done.record(side)
                dones.append(done)
        self._ops.heff_apply_gated(self.bufs[b], L_rows, M1, M2, R, out, self.flags, self.world, self.rank, self.epoch, 1,
                                   li, ri, lrow0)
        for done in dones:
            main.wait_event(done)        # v_shard may be reused once our outgoing copies have left
        return out

    # ---- host-resident call: theta shard in pinned host memory, result rows to pinned host memory ---------------
    def forward_host(self, v_host: torch.Tensor, L_rows, M1, M2, R, out_host: torch.Tensor) -> torch.Tensor:
        """The same sharded matvec with this rank's shard of theta in PINNED HOST memory and its rows of the result
        delivered to pinned host memory, the copies pipelined with the compute:

          * the shard is uploaded straight into this rank's slot of the gathered buffer in COLUMN blocks
            (``qsb_upload_columns``, copy engine, one flag per block);
          * two side streams forward every block to the peers as soon as it has landed
            (``qsb_push_columns_dma``: stream-ordered wait on the upload's flag, 2-D peer copy, flag write on the peer);
          * the step-1 GEMM is gated on the (source rank x column block) flag grid (``chunk_axis = 2``): tiles of column
            block j start when block j of the row block they read is there, so the math begins after 1/grid_cols of the
            transfers instead of after all of them;
          * the result is produced in row blocks (1/2, 1/4, ... of the rows, >= 128 each); block i travels to the host
            while block i + 1 is computed.

        ASYNCHRONOUS like ``ApplyMPO``'s host path: ``out_host`` is complete (and ``v_host`` reusable) once the launching
        stream has reached the end of this call -- synchronise the stream or ``self.host_done``."""
        if not (v_host.device.type == "cpu" and v_host.is_pinned() and out_host.device.type == "cpu"
                and out_host.is_pinned()):
            raise ValueError("forward_host: theta shard and result must be pinned host tensors")
        ar = self.arena
        ar.epoch += 1
        ep, b = ar.epoch, ar.epoch & 1
        dev = L_rows.device
        main = torch.cuda.current_stream(dev)
        h2d, d2h = ar.copy_streams["h2d"], ar.copy_streams["d2h"]
        nloc = self.chi_l * self.per_row // self.world
        rows = self.hi - self.lo
        # column blocks: as many as the flag grid allows, each a whole number of GEMM tile columns (BN = 128 float64 /
        # 64 complex128) -- every rank derives the same count from the same shape; 1 block = plain per-source gating
        nc = ar.grid_cols
        bn = 64 if v_host.dtype.is_complex else 128
        while nc > 1 and (self.per_row % nc != 0 or (self.per_row // nc) % bn != 0):
            nc //= 2
        own = self.bufs[b][self.rank * nloc:(self.rank + 1) * nloc]
        own_flags = ar.grid_flags[self.rank * nc:(self.rank + 1) * nc]
        peer_flag_ptrs = [ar.ptrs[r] + 4 * (32 + self.rank * nc) for r in range(self.world)]
        start = torch.cuda.Event()
        start.record(main)               # everything this buffer parity was used for two steps ago is behind `main`
        h2d.wait_event(start)
        with torch.cuda.stream(h2d):
            self._lib.upload_columns(v_host.reshape(-1), own, rows, self.per_row, nc, own_flags, ep)
        peers = self._order[1:]
        for si, side in enumerate(ar.side):
            targets = peers[si::len(ar.side)]
            if not targets:
                continue
            side.wait_event(start)
            with torch.cuda.stream(side):
                self._lib.push_columns_dma(own, rows, self.per_row, nc, own_flags,
                                           [self.dst_ptrs[b][r] for r in targets],
                                           [peer_flag_ptrs[r] for r in targets], ep)
        (li, lrow0), (ri, _r0) = self._lib.identity_of(L_rows), self._lib.identity_of(R)
        per_out = M1.shape[2] * M2.shape[2] * R.shape[1]
        out_dev = self._host_out(rows * per_out, v_host.dtype, dev)
        bounds = self._lib.row_block_bounds(rows)          # row blocks of the result, big first, small last
        for r_lo, r_hi in zip(bounds[:-1], bounds[1:]):
            Lb = L_rows[:, r_lo:r_hi, :]
            ob = out_dev[r_lo * per_out:r_hi * per_out]
            self._ops.heff_apply_gridgated(self.bufs[b], Lb, M1, M2, R, ob, ar.grid_flags, self.world, nc, self.rank, ep,
                                           li, ri, lrow0 + r_lo if li >= 0 else 0)
            done = torch.cuda.Event()
            done.record(main)
            d2h.wait_event(done)
            with torch.cuda.stream(d2h):
                out_host[r_lo * per_out:r_hi * per_out].copy_(ob, non_blocking=True)
        for st in (h2d, d2h) + tuple(ar.side):
            main.wait_stream(st)
        self.host_done = torch.cuda.Event()
        self.host_done.record(main)
        return out_host

    def _host_out(self, n: int, dtype, dev) -> torch.Tensor:
        buf = getattr(self, "_host_out_buf", None)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(n, dtype=dtype, device=dev)
            self._host_out_buf = buf
        return buf[:n]


def lanczos_ground_state_sharded(initial_shard: torch.Tensor, linear_operator: Callable,
                                 operator_args: Sequence = (), num_restarts: int = 2, krylov_dim: int = 4, *,
                                 pro: bool = False, pro_max_reorth: int = 1, beta_tol: float = 1e-12,
                                 ritz_residual_tol: Optional[float] = None, group: Any = None,
                                 _vector_ops=None) -> Tuple[torch.Tensor, float]:
    """``lanczos_ground_state`` on row shards: every rank passes its shard of the start vector and a
    ``linear_operator`` that maps shard -> shard (e.g. ``ShardedApplyMPO``); returns this rank's shard of the
    ground-state vector and the (identical on all ranks) Ritz value."""
    import torch.distributed as dist
    if group is None:
        group = dist.group.WORLD
    vec = _vector_ops if _vector_ops is not None else _CUDA_OPS
    if _vector_ops is None and initial_shard.device.type != "cuda":
        raise RuntimeError("qsb200 sharded Lanczos runs on CUDA tensors only (no CPU fallback)")
    return _lanczos_impl(initial_shard, linear_operator, tuple(operator_args), num_restarts, krylov_dim, pro,
                         pro_max_reorth, beta_tol, ritz_residual_tol, group=group, vec=vec)


# =====================================================================================================
# Sharded environments and the sharded two-site DMRG driver
# =====================================================================================================
class CudaEngine:
    """The compute engine of the sharded driver: this package's CUDA kernels.  (The gloo tests of the multi-rank
    HOST logic inject a CPU engine with the same methods built on the oracle; the product never does.)"""

    def __init__(self):
        from .dmrg import ApplyMPO, LeftEnvUpdate, RightEnvUpdate
        from . import _lib
        self._lib = _lib
        self.apply_mpo = ApplyMPO()
        self._left, self._right = LeftEnvUpdate(), RightEnvUpdate()
        self.vector_ops = None                       # None -> CudaVectorOps

    fused_scale = True          # matmul takes row_scale / col_scale (Schmidt weights in the GEMM epilogue)

    def matmul(self, a, b, row_scale=None, col_scale=None):
        from .dmrg import _matmul2d
        return _matmul2d(a, b, row_scale=row_scale, col_scale=col_scale)

    def env_left(self, Lp, M, A):
        return self._left(Lp, M, A)

    def env_right(self, M, Rn, B):
        return self._right(M, Rn, B)

    def env_left_rows(self, Lp_rows, M, A, A_rows):
        """Partial sum of the full new left environment from the bra-row block this rank owns."""
        out = torch.empty(M.shape[1], A.shape[2], A.shape[2], device=Lp_rows.device, dtype=Lp_rows.dtype)
        self._lib.ops.env_left_rows(Lp_rows, M.contiguous(), A, A_rows, out)
        return out

    def env_right_rows(self, M, Rn, B, B_rows):
        """Rows [lo, hi) of the new right environment (no communication needed)."""
        out = torch.empty(M.shape[0], B_rows.shape[0], B.shape[0], device=Rn.device, dtype=Rn.dtype)
        self._lib.ops.env_right_rows(Rn, M.contiguous(), B, B_rows, out)
        return out

    def split(self, theta2d, maxdim, cutoff, **kw):
        from .svd import two_site_split
        return two_site_split(theta2d, maxdim, cutoff, **kw)

    def tag(self, env, row0=0):
        """Test an environment (or this rank's row block of one, first row ``row0``) for an identity channel and
        record the answer on the tensor: the matvec kernels skip the chi^3 GEMM slice of such a channel."""
        if self._lib.identity_of(env)[0] < 0:
            self._lib.tag_identity(env, row0)
        return env

    def lanczos(self, v, op, args, **kw):
        from .eigensolver import lanczos_ground_state
        return lanczos_ground_state(v, op, args, **kw)


def _real_view(t: torch.Tensor) -> torch.Tensor:
    """NCCL reductions take real dtypes only; a complex tensor is reduced as its (re, im) float64 view."""
    return torch.view_as_real(t) if t.is_complex() else t


class _Env:
    """One cached environment: either replicated (full (W, chi, chi)) or this rank's row block (W, rows, chi)."""
    __slots__ = ("t", "sharded", "chi")

    def __init__(self, t: torch.Tensor, sharded: bool, chi: int):
        self.t, self.sharded, self.chi = t, sharded, chi


class _SiteStore:
    """The site tensors of one canonical form (``mps.A`` or ``mps.B``), held either replicated (the reference's layout)
    or -- ``sharded=True`` -- as this rank's ROW BLOCK over the left bond wherever that bond is shardable.  At
    chi = 4096, d = 4, complex128 one site tensor is 1 GB and the 2 N replicated tensors of a 64-site chain would be
    137 GB per GPU; row-sharded they are 17 GB, and a sweep only ever needs one full tensor (the right factor of theta /
    the ket of an environment update) at a time, which is all-gathered on demand."""

    def __init__(self, plist, owner: "ShardedDMRG"):
        self.plist, self.o = plist, owner
        self.left_dim = {}                  # p -> true left bond dimension of a tensor stored as a row block
        self._cache = (None, None)          # (p, full tensor): the factor the last split produced is reused once

    def is_sharded(self, p: int) -> bool:
        return p in self.left_dim

    def rows(self, p: int, lo: int, hi: int) -> torch.Tensor:
        """Rows [lo, hi) of the left bond of site p -- this rank's own block when the tensor is stored sharded."""
        t = self.plist[p].detach()
        if self.is_sharded(p):
            assert (lo, hi) == self.o._block(self.left_dim[p]), "a sharded site tensor only serves its own row block"
            return t
        return t[lo:hi]

    def shape(self, p: int):
        t = self.plist[p]
        return (self.left_dim.get(p, t.shape[0]), t.shape[1], t.shape[2])

    def full(self, p: int) -> torch.Tensor:
        import torch.distributed as dist
        t = self.plist[p].detach()
        if not self.is_sharded(p):
            return t
        if self._cache[0] == p:
            return self._cache[1]
        l = self.left_dim[p]
        out = torch.empty(l, t.shape[1], t.shape[2], dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.o.group)
        self.o.comm["all_gather_site"] = self.o.comm.get("all_gather_site", 0) + 1
        return out

    def set(self, p: int, full: torch.Tensor, shard: bool) -> None:
        import torch.nn as nn
        l = full.shape[0]
        if shard and self.o._shardable(l):
            lo, hi = self.o._block(l)
            self.plist[p] = nn.Parameter(full[lo:hi].contiguous(), requires_grad=False)
            self.left_dim[p] = l
            self._cache = (p, full)
        else:
            self.plist[p] = nn.Parameter(full, requires_grad=False)
            self.left_dim.pop(p, None)
            if self._cache[0] == p:
                self._cache = (None, None)

    def drop_cache(self) -> None:
        self._cache = (None, None)


class ShardedDMRG:
    """Two-site DMRG with the inner loop sharded over the left bond index across the ranks of ``group``.

    Same schedule semantics and ``history`` keys as ``DMRG`` (reference dmrg.py:506-802).  Every rank holds the
    (small) MPS replicated and runs the same control flow.  What is distributed:
      * environments ``L[p]`` / ``R[p]`` are stored as row blocks over their bra index whenever the bond is
        shardable (``chi % world == 0 and chi >= min_shard_chi``), else replicated;
      * the Lanczos solve runs on row shards (``lanczos_ground_state_sharded`` + ``ShardedApplyMPO``): one
        all-gather of the vector per matvec, one small all-reduce per scalar;
      * ``update_L`` contracts the sharded bra index: every rank produces a partial sum of the new environment,
        combined and re-sharded with one reduce-scatter; ``update_R`` computes its own row block locally;
      * the right environment of the active bond is all-gathered once per bond step (it multiplies the unsharded
        right bond of theta);
      * theta is built sharded (every rank forms its rows) and the two-site split of truncating bulk blocks runs ON the
        row shards (``svd.sketch_split_sharded``: local GEMMs, l x l all-reduces, the small eigh on rank 0 and its
        result broadcast); other blocks are gathered, split on rank 0 and the factors broadcast -- either way all ranks
        hold bit-identical MPS tensors (a replicated cuSOLVER call is not guaranteed to be bit-reproducible across
        devices);
      * noise (dmrg.py:677-682) is drawn on rank 0 and broadcast.
    """

    def __init__(self, mps, mpo, *, device, dtype=torch.float64, group=None, min_shard_chi: int = 64, engine=None,
                 use_fused: bool = True, shard_mps: bool = False):
        import torch.distributed as dist
        from .dmrg import _as_mpo_list, _check_mpo_shapes
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device, self.dtype = device, dtype
        self.mps = mps
        self.Nsites = len(mpo)
        self.Ms = _as_mpo_list(list(mpo), self.Nsites, device, dtype)
        self.chid = self.Ms[0].shape[2]
        ML = torch.ones(self.Ms[0].shape[0], 1, 1, device=device, dtype=dtype)
        MR = torch.ones(self.Ms[-1].shape[1], 1, 1, device=device, dtype=dtype)
        _check_mpo_shapes(self.Ms, ML, MR, self.chid)
        self.min_shard_chi = int(min_shard_chi)
        self.engine = engine if engine is not None else CudaEngine()
        self.history = self._empty_history()
        self.use_fused = bool(use_fused)
        self._ops_cache: dict = {}
        self._arena: Optional[SymmetricArena] = None
        self._max_chi = 0
        self.sharded_split = True       # two-site split on the row shards where the block allows it (svd.py)
        # shard_mps: keep the MPS itself row-sharded over the left bonds (see _SiteStore); `mps.A` / `mps.B` then hold
        # this rank's blocks until gather_mps() reassembles them
        self.shard_mps = bool(shard_mps)
        self._A, self._B = _SiteStore(mps.A, self), _SiteStore(mps.B, self)
        prev = getattr(mps, "_qsb_row_sharded", None)
        if prev is not None:            # an MPS left sharded by an earlier ShardedDMRG on the same group
            self._A.left_dim, self._B.left_dim = dict(prev[0]), dict(prev[1])
        self.profile = False            # True: synchronise around the phases and accumulate seconds in self.timers
        self.timers = {"theta": 0.0, "gather_R": 0.0, "lanczos": 0.0, "split": 0.0, "env": 0.0}
        self.comm = {"all_gather_vec": 0, "all_gather_env": 0, "reduce_scatter_env": 0, "all_reduce_env": 0,
                     "broadcast_split": 0, "sharded_steps": 0, "replicated_steps": 0}

    @staticmethod
    def _empty_history() -> dict:
        return {"energies": [], "discarded_weights": [], "bond_dims": [], "sweeps": [], "directions": [], "sites": []}

    # ---- sharding policy / conversions ------------------------------------------------------------------
    def _shardable(self, chi: int) -> bool:
        return self.world > 1 and chi >= self.min_shard_chi and chi % self.world == 0

    def _block(self, chi: int) -> Tuple[int, int]:
        return row_partition(chi, self.world)[self.rank]

    def _full(self, e: _Env) -> torch.Tensor:
        import torch.distributed as dist
        if not e.sharded:
            return e.t
        W, rows, chi = e.t.shape
        buf = torch.empty(self.world * W * rows * chi, dtype=e.t.dtype, device=e.t.device)
        dist.all_gather_into_tensor(buf, e.t.contiguous().view(-1), group=self.group)
        self.comm["all_gather_env"] += 1
        return buf.view(self.world, W, rows, chi).permute(1, 0, 2, 3).reshape(W, self.world * rows, chi)

    def _rows(self, e: _Env) -> torch.Tensor:
        if e.sharded:
            return e.t
        lo, hi = self._block(e.chi)
        return e.t[:, lo:hi, :]

    def _tag(self, t: torch.Tensor, row0: int = 0) -> torch.Tensor:
        tag = getattr(self.engine, "tag", None)           # the CPU engines of the gloo tests have no such shortcut
        return tag(t, row0) if tag is not None else t

    def _store(self, full: torch.Tensor) -> _Env:
        chi = full.shape[1]
        if self._shardable(chi):
            lo, hi = self._block(chi)
            return _Env(self._tag(full[:, lo:hi, :].contiguous(), lo), True, chi)
        return _Env(self._tag(full), False, chi)

    # ---- environment updates ---------------------------------------------------------------------------
    def _update_L(self, Lp: _Env, M: torch.Tensor, A: torch.Tensor) -> _Env:
        import torch.distributed as dist
        chi_new = A.shape[2]
        if not Lp.sharded:
            return self._store(self.engine.env_left(Lp.t, M, A))
        lo, hi = self._block(Lp.chi)
        partial = self.engine.env_left_rows(Lp.t, M, A, A[lo:hi])          # (W', chi', chi'), partial sum
        if self._shardable(chi_new):
            Wn = partial.shape[0]
            rows = chi_new // self.world
            send = partial.permute(1, 0, 2).contiguous()                    # rows outermost for the scatter
            recv = torch.empty(rows * Wn * chi_new, dtype=partial.dtype, device=partial.device)
            if dist.get_backend(self.group) == "gloo":    # gloo (CPU tests) has no reduce-scatter: all-reduce + slice
                flat = send.view(-1)
                dist.all_reduce(_real_view(flat), group=self.group)
                recv.copy_(flat[self.rank * recv.numel(): (self.rank + 1) * recv.numel()])
            else:                                         # NCCL reduces real dtypes only: (re, im) pairs
                dist.reduce_scatter_tensor(_real_view(recv), _real_view(send.view(-1)), group=self.group)
            self.comm["reduce_scatter_env"] += 1
            new = recv.view(rows, Wn, chi_new).permute(1, 0, 2)                 # strided view, last dim contiguous
            return _Env(self._tag(new, self._block(chi_new)[0]), True, chi_new)
        dist.all_reduce(_real_view(partial), group=self.group)
        self.comm["all_reduce_env"] += 1
        return _Env(self._tag(partial), False, chi_new)

    def _update_R(self, M: torch.Tensor, Rn_full: torch.Tensor, B: torch.Tensor) -> _Env:
        chi_new = B.shape[0]
        if self._shardable(chi_new):
            lo, hi = self._block(chi_new)
            return _Env(self.engine.env_right_rows(M, Rn_full, B, B[lo:hi]), True, chi_new)
        return _Env(self.engine.env_right(M, Rn_full, B), False, chi_new)

    # ---- one bond step ----------------------------------------------------------------------------------
    def _ground_state(self, theta: torch.Tensor, L: _Env, M1, M2, R_full, chi_l: int, kw: dict):
        """Lanczos on this bond.  Sharded bond: ``theta`` is this rank's row shard (flat) and so is the result;
        replicated bond: full vectors."""
        per_row = M1.shape[3] * M2.shape[3] * R_full.shape[2]
        if L.sharded:
            op = self._sharded_op(chi_l, per_row, theta.dtype, theta.device)
            v, E = lanczos_ground_state_sharded(theta.reshape(-1), op, (L.t, M1, M2, R_full), group=self.group,
                                                _vector_ops=self.engine.vector_ops, **kw)
            self.comm["all_gather_vec"] += 1 + kw["num_restarts"] * kw["krylov_dim"]
            self.comm["sharded_steps"] += 1
            return v, E
        self.comm["replicated_steps"] += 1
        v, E = self.engine.lanczos(theta.reshape(-1), self.engine.apply_mpo, (L.t, M1, M2, R_full), **kw)
        return v, E

    def _sharded_op(self, chi_l: int, per_row: int, dtype, device):
        """The sharded matvec for this bond: the fused P2P-push + gated-GEMM operator when the bond allows it (CUDA
        engine, chi_l divisible by BK * world), else all-gather + local matvec.  Cached per shape: building the fused
        operator allocates symmetric memory and is a collective, which every rank reaches in the same order."""
        key = (chi_l, per_row, dtype)
        op = self._ops_cache.get(key)
        if op is None:
            if self.use_fused and isinstance(self.engine, CudaEngine) \
                    and fused_eligible(chi_l, per_row, dtype, self.world):
                es = torch.empty(0, dtype=dtype).element_size()
                if self._arena is None or self._arena.stride < chi_l * per_row * es:
                    # one symmetric allocation for the largest vector this run can see (maxdim^2 d^2)
                    big = max(chi_l * per_row, self._max_chi * self._max_chi * self.chid * self.chid)
                    self._arena = SymmetricArena(big * es, device, self.group)
                    self.comm["arenas_built"] = self.comm.get("arenas_built", 0) + 1
                op = FusedShardedApplyMPO(chi_l, per_row, dtype, device, group=self.group, arena=self._arena)
                self.comm["fused_ops_built"] = self.comm.get("fused_ops_built", 0) + 1
            else:
                op = ShardedApplyMPO(chi_l, per_row, group=self.group, local_op=self.engine.apply_mpo)
            self._ops_cache[key] = op
        return op

    def _add_noise(self, v: torch.Tensor, noise: float, sharded: bool, chi_l: int) -> torch.Tensor:
        """The reference's noise term (dmrg.py:677-682: uniform [0,1) vector from the global RNG, normalised, added with
        weight ``noise``, result renormalised) on a replicated vector or on this rank's row shard.  Rank 0 draws the
        vector (its global RNG plays the role of the reference's) and broadcasts it, so every rank perturbs the same
        state whatever the RNG states of the other ranks are."""
        import torch.distributed as dist
        n_full = v.numel() * (self.world if sharded else 1)
        nv = torch.empty(n_full, dtype=v.dtype, device=v.device)
        if self.rank == 0:
            nv.copy_(torch.rand_like(nv))
        dist.broadcast(_real_view(nv), src=dist.get_global_rank(self.group, 0), group=self.group)
        self.comm["broadcast_noise"] = self.comm.get("broadcast_noise", 0) + 1
        nv.div_(torch.linalg.norm(nv))
        if sharded:
            part = nv.view(self.world, -1)[self.rank]
            out = v + part * noise
            n2 = (torch.linalg.norm(out) ** 2).reshape(1)
            dist.all_reduce(n2, group=self.group)
            return out.div_(torch.sqrt(n2).to(out.dtype))
        out = v + nv * noise
        return out.div_(torch.linalg.norm(out))

    def _split_any(self, v: torch.Tensor, sharded: bool, chil: int, chir: int, chi: int, cutoff: float):
        """Two-site split of the Lanczos result: on the row shards where the block allows it
        (``svd.sketch_split_sharded``: nobody gathers theta, only l x l matrices are reduced), else gathered and split on
        rank 0."""
        from . import svd as _svd
        from .svd import sharded_split_eligible, sketch_split_sharded, gram_split_sharded
        FAST_MIN = _svd.FAST_MIN
        d = self.chid
        m, n = chil * d, d * chir
        tried = False
        if sharded and self.sharded_split and sharded_split_eligible(m, n, chi, self.world):
            tried = True
            out = sketch_split_sharded(v.view(m // self.world, n), m, chi, cutoff, self.group, self.comm)
            if out is not None:
                self.comm["sharded_splits"] = self.comm.get("sharded_splits", 0) + 1
                return out
        if sharded and self.sharded_split and n <= m and m % self.world == 0 and min(m, n) >= FAST_MIN:
            # real truncation (or no truncation at all): the exact Gram/eigh split, still without gathering theta
            out = gram_split_sharded(v.view(m // self.world, n), m, chi, cutoff, self.group, self.comm)
            if out is not None:
                self.comm["sharded_gram_splits"] = self.comm.get("sharded_gram_splits", 0) + 1
                return out
        theta = gather_rows(v, chil, d * d * chir, self.group) if sharded else v
        return self._split(theta.reshape(m, n), chi, cutoff, sketch=False if tried else None)

    def _split(self, theta2d: torch.Tensor, chi: int, cutoff: float, sketch=None):
        """Two-site split on rank 0, factors broadcast to everyone."""
        import torch.distributed as dist
        m, n = theta2d.shape
        src = dist.get_global_rank(self.group, 0)
        meta = torch.zeros(2, dtype=torch.int64, device=theta2d.device)
        if self.rank == 0:
            kw = {} if sketch is None else {"sketch": sketch}
            U, S, Vh, keep = self.engine.split(theta2d, chi, cutoff, **kw)
            meta[0], meta[1] = keep, S.numel()
        dist.broadcast(meta, src=src, group=self.group)
        keep, ns = int(meta[0]), int(meta[1])
        rdt = torch.float64 if theta2d.dtype in (torch.float64, torch.complex128) else torch.float32
        if self.rank != 0:
            U = torch.empty(m, keep, dtype=theta2d.dtype, device=theta2d.device)
            S = torch.empty(ns, dtype=rdt, device=theta2d.device)
            Vh = torch.empty(keep, n, dtype=theta2d.dtype, device=theta2d.device)
        else:
            U, S, Vh = U.resolve_conj().contiguous(), S.contiguous(), Vh.resolve_conj().contiguous()
        for t in (U, S, Vh):
            dist.broadcast(t, src=src, group=self.group)
        self.comm["broadcast_split"] += 1
        return U, S, Vh, keep

    def _sweep(self, direction: str, L: list, R: list, params: dict, Ekeep: list) -> None:
        import torch.nn as nn
        mps, d, N = self.mps, self.chid, self.Nsites
        sW = mps.sWeight
        sites = range(N - 1) if direction == "right" else range(N - 2, -1, -1)
        kw = dict(num_restarts=params["maxit"], krylov_dim=params["krydim"], pro=params["reortho"],
                  pro_max_reorth=params["maxreortho"])
        if self.profile:
            import time
            torch.cuda.synchronize()
            self._t_last = time.perf_counter()
        for p in sites:
            # theta build (dmrg.py:633-660).  On a sharded bond every rank builds only ITS rows of theta (the row block
            # of the left site tensor times the full right one): the sharded Lanczos solver takes exactly that shard
            st = self._B if direction == "right" else self._A
            sw = sW[p] if direction == "right" else sW[p + 2]
            if sw.dim() == 2:
                sw = torch.diagonal(sw, 0)
            l, s, m = st.shape(p)
            rt = st.full(p + 1)                                  # the right factor is needed whole (one all-gather)
            _m, t, r = rt.shape
            chil, chir = l, r
            sharded = L[p].sharded
            lo, hi = self._block(chil) if sharded else (0, l)
            from .dmrg import build_theta
            psi = build_theta(st.rows(p, lo, hi), rt, sw[lo:hi] if direction == "right" else sw, direction,
                              matmul=self.engine.matmul, fused=getattr(self.engine, "fused_scale", False))
            del rt
            st.drop_cache()
            self._tick("theta")
            R_full = self._tag(self._full(R[p + 2]))
            self._tick("gather_R")
            theta, E = self._ground_state(psi, L[p], self.Ms[p], self.Ms[p + 1], R_full, chil, kw)
            self._tick("lanczos")
            Ekeep.append(E)
            self.history["energies"].append(float(E))
            if params["noise"] > 0:
                theta = self._add_noise(theta, params["noise"], sharded, chil)
            Uk, S, Vhk, keep = self._split_any(theta, sharded, chil, chir, params["chi"], params["cutoff"])
            self._tick("split")
            self.history["discarded_weights"].append(torch.sum(S[keep:] ** 2).real.item())
            self.history["bond_dims"].append(int(keep))
            self.history["sweeps"].append(int(params["sweep_idx"]))
            self.history["directions"].append(direction)
            self.history["sites"].append(int(p))
            sv = S[:keep]
            denom = sv.norm().add_(torch.finfo(sv.dtype).eps)
            A_new, B_new = Uk.reshape(chil, d, keep), Vhk.reshape(keep, d, chir)
            sW[p + 1] = nn.Parameter(sv.div(denom), requires_grad=False)
            if direction == "right":
                L[p + 1] = self._update_L(L[p], self.Ms[p], A_new)
            else:
                R[p + 1] = self._update_R(self.Ms[p + 1], R_full, B_new)
            self._A.set(p, A_new, self.shard_mps)               # stored whole or as this rank's row block
            self._B.set(p + 1, B_new, self.shard_mps)
            del A_new, B_new, Uk, Vhk
            self._tick("env")
            if params["dispon"] == 2 and self.rank == 0:
                print(f"Sweep: {params['sweep_idx']} of {params['numsweeps']}, Loc: {p}, Energy: {Ekeep[-1]:.6f}")

    def _tick(self, phase: str) -> None:
        if not self.profile:
            return
        import time
        torch.cuda.synchronize()
        now = time.perf_counter()
        self.timers[phase] += now - self._t_last
        self._t_last = now

    def _boundary(self, mat: torch.Tensor):
        """Small boundary SVD (dmrg.py:776, 793): rank 0 computes, everyone receives the same factors."""
        import torch.distributed as dist
        U, S, Vh = torch.linalg.svd(mat, full_matrices=False)
        U, S, Vh = U.resolve_conj().contiguous(), S.contiguous(), Vh.resolve_conj().contiguous()
        src = dist.get_global_rank(self.group, 0)
        for t in (U, S, Vh):
            dist.broadcast(t, src=src, group=self.group)
        return U, S, Vh

    def forward(self, sweeps, *, dispon=2, updateon=True, maxit=2):
        import torch.nn as nn
        if not updateon:
            raise ValueError("ShardedDMRG always updates (updateon=False is a single-GPU debugging mode)")
        N, d = self.Nsites, self.chid
        Ekeep: list = []
        self.history = self._empty_history()
        self._max_chi = max(int(x) for x in sweeps.maxdim)
        L: list = [None] * (N + 1)
        R: list = [None] * (N + 1)
        L[0] = _Env(torch.ones(self.Ms[0].shape[0], 1, 1, device=self.device, dtype=self.dtype), False, 1)
        R[N] = _Env(torch.ones(self.Ms[-1].shape[1], 1, 1, device=self.device, dtype=self.dtype), False, 1)
        for p in range(N - 1, self.mps.ortho_center, -1):                          # dmrg.py:752-753
            R[p] = self._update_R(self.Ms[p], self._full(R[p + 1]), self._B.full(p))
        if self.shard_mps:
            self._shard_now()
        for k in range(1, sweeps.numsweeps + 1):
            params = {"chi": sweeps.maxdim[k - 1], "krydim": sweeps.krylov_dim[k - 1], "cutoff": sweeps.cutoff[k - 1],
                      "reortho": sweeps.reortho[k - 1], "maxreortho": sweeps.maxreortho[k - 1], "dispon": dispon,
                      "maxit": maxit, "sweep_idx": k, "numsweeps": sweeps.numsweeps, "noise": sweeps.noise[k - 1]}
            self._sweep("right", L, R, params, Ekeep)
            self.mps.ortho_center = N - 1
            sw = self.mps.sWeight[N - 1]                                            # dmrg.py:769-781
            if sw.dim() == 2:
                sw = torch.diagonal(sw, 0)
            Bl = self._B.full(N - 1)
            U, S, Vh = self._boundary((sw.to(Bl.dtype).view(-1, 1, 1) * Bl).reshape(Bl.shape[0] * d, Bl.shape[2]))
            eps = torch.finfo(S.dtype).eps
            self._A.set(N - 1, U.view(Bl.shape[0], d, Bl.shape[2]), self.shard_mps)
            self.mps.sWeight[N] = nn.Parameter((S.unsqueeze(1) * Vh).to(self.dtype) / (S.norm().add_(eps)),
                                               requires_grad=False)
            self._sweep("left", L, R, params, Ekeep)
            self.mps.ortho_center = 0
            sw = self.mps.sWeight[1]                                                # dmrg.py:787-796
            if sw.dim() == 2:
                sw = torch.diagonal(sw, 0)
            A0f = self._A.full(0)
            A0 = A0f * sw.to(A0f.dtype).view(1, 1, -1)
            chil, _, chir = A0f.shape
            U, S, Vh = self._boundary(A0.reshape(chil, d * chir))
            self._B.set(0, Vh.view(chil, d, chir), self.shard_mps)
            self.mps.sWeight[0] = nn.Parameter((U * S.unsqueeze(0)).to(self.dtype) / (S.norm().add_(eps)),
                                               requires_grad=False)
            if dispon == 1 and self.rank == 0:
                print(f"Sweep: {k} of {sweeps.numsweeps}, Energy: {Ekeep[-1]:.12f}, "
                      f"Bond dim: {self.mps.A[N // 2].shape[-1]}")
        self.mps._qsb_row_sharded = (dict(self._A.left_dim), dict(self._B.left_dim)) if self.shard_mps else None
        return Ekeep, self.mps

    __call__ = forward

    def _shard_now(self) -> None:
        """Replace every replicated site tensor whose left bond is shardable by this rank's row block."""
        for store in (self._A, self._B):
            for p in range(self.Nsites):
                if not store.is_sharded(p):
                    store.set(p, store.plist[p].detach(), True)
            store.drop_cache()
        if self.device is not None and torch.device(self.device).type == "cuda":
            torch.cuda.empty_cache()

    def gather_mps(self):
        """Reassemble a row-sharded MPS (``shard_mps=True``) into the reference's replicated layout on every rank."""
        for store in (self._A, self._B):
            for p in range(self.Nsites):
                if store.is_sharded(p):
                    store.set(p, store.full(p), False)
            store.drop_cache()
        self.mps._qsb_row_sharded = None
        return self.mps
