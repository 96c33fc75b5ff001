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

__all__ = ["row_partition", "ShardedApplyMPO", "lanczos_ground_state_sharded", "shard_rows", "gather_rows"]


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
