"""Two-site DMRG on the sm_100a kernels: host-side mirror of ``quantum_simulation/dmrg.py``.

Same public names, arguments and error behaviour as the reference (file:line relative to
/root/reference/quantum_simulation/):

  ApplyMPO            dmrg.py:165-248     H_eff . theta              -> torch.ops.qsb200.heff_apply
  LeftEnvUpdate       dmrg.py:251-278     L[p+1] from L[p], M[p], A  -> torch.ops.qsb200.env_left
  RightEnvUpdate      dmrg.py:281-308     R[p]   from M[p], R[p+1], B-> torch.ops.qsb200.env_right
  EnvironmentManager  dmrg.py:311-418
  Sweeps              dmrg.py:424-503
  DMRG                dmrg.py:506-802
  set_contract_backend dmrg.py:85-107

The ``contract_backend`` names ``'auto' | 'einsum' | 'ncon'`` are accepted so notebooks run unchanged, but all
three route to the same CUDA kernels -- there is no multi-backend dispatch and no CPU fallback.  Results follow
the reference's ncon path (the optimal contraction order) to ~1e-15 relative.

Environment layout: the reference's ncon branch returns environments as a permuted view with strides
``(chi, 1, W*chi)``; ours are contiguous ``(W, chi, chi)`` -- the kernels take either.
"""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .eigensolver import lanczos_ground_state
from .mps import MPS
from .svd import two_site_split

__all__ = ["ApplyMPO", "LeftEnvUpdate", "RightEnvUpdate", "EnvironmentManager", "Sweeps", "DMRG",
           "set_contract_backend"]

_BACKENDS = ("auto", "einsum", "ncon")
_CONTRACT_BACKEND = "qsb200"


def set_contract_backend(backend: str = "auto", device: Any = "cpu") -> None:
    """Validates the reference's backend name (dmrg.py:85-107).  Every valid name resolves to the CUDA kernels."""
    dev = device if isinstance(device, str) else getattr(device, "type", "cpu")
    if backend not in _BACKENDS:
        raise ValueError(f"Unknown backend: {backend!r} (valid: 'auto'|'einsum'|'ncon')")
    print(f"[DMRG] Contraction backend resolved to 'qsb200-cuda' (requested '{backend}') on device='{dev}'.")


def _as_mpo_list(M_or_Ms: Any, Nsites: int, device: Any, dtype: torch.dtype) -> List[torch.Tensor]:
    """dmrg.py:17-43 (numpy site tensors are accepted here; the reference's numpy branch cannot run)."""
    def conv(m):
        if isinstance(m, np.ndarray):
            m = torch.from_numpy(np.ascontiguousarray(m))
        if m.is_complex() and not dtype.is_complex:
            if float(m.imag.abs().max()) != 0.0:
                raise ValueError("MPO has a non-zero imaginary part but a real dtype was requested "
                                 "(the reference silently drops it, giving the wrong Hamiltonian)")
            m = m.real
        return m.to(device=device, dtype=dtype).contiguous()
    if isinstance(M_or_Ms, (list, tuple)):
        assert len(M_or_Ms) == Nsites, f"MPO list length {len(M_or_Ms)} != Nsites {Nsites}"
        return [conv(m) for m in M_or_Ms]
    return [conv(M_or_Ms) for _ in range(Nsites)]


def _check_mpo_shapes(Ms: List[torch.Tensor], ML: torch.Tensor, MR: torch.Tensor, chid: int) -> None:
    """dmrg.py:46-77."""
    for p, M in enumerate(Ms):
        assert M.ndim == 4, f"Ms[{p}] must be rank-4 (Wl,Wr,d,d), got {tuple(M.shape)}"
        assert M.shape[2] == chid and M.shape[3] == chid, \
            f"Ms[{p}] physical dims {tuple(M.shape[2:])} != (d,d)=({chid},{chid})"
        if p < len(Ms) - 1:
            assert M.shape[1] == Ms[p + 1].shape[0], \
                f"MPO bond mismatch: Ms[{p}].Wr={M.shape[1]} != Ms[{p + 1}].Wl={Ms[p + 1].shape[0]}"
    assert ML.shape[0] == Ms[0].shape[0], f"ML W={ML.shape[0]} must match Ms[0].Wl={Ms[0].shape[0]}"
    assert MR.shape[0] == Ms[-1].shape[1], f"MR W={MR.shape[0]} must match Ms[-1].Wr={Ms[-1].shape[1]}"


def _matmul2d(a: torch.Tensor, b: torch.Tensor, row_scale=None, col_scale=None) -> torch.Tensor:
    """(m x k) @ (k x n) on the library's FP64 tensor-core GEMM core (theta build, dmrg.py:645-646 / 659-660);
    operands with no unit stride are made contiguous first (they are chi^2 d elements).  ``row_scale`` / ``col_scale``:
    (real vector, divisor / modulus) fused into the GEMM epilogue (``_lib.gemm``)."""
    if a.stride(0) != 1 and a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(0) != 1 and b.stride(1) != 1:
        b = b.contiguous()
    return _lib.gemm(_dev_tensor(a), _dev_tensor(b), row_scale=row_scale, col_scale=col_scale)


def build_theta(left_t: torch.Tensor, right_t: torch.Tensor, sw: torch.Tensor, direction: str, matmul=_matmul2d,
                fused: bool = True) -> torch.Tensor:
    """The two-site wavefunction of a bond (dmrg.py:633-662): ``(sw . left) @ right`` in a right sweep (``sw`` on the
    left bond of ``left_t``), ``left @ (right . sw)`` in a left sweep (``sw`` on the right bond of ``right_t``), as ONE
    GEMM with the Schmidt weights applied to the rows (l, s) -> sw[l] resp. columns (t, r) -> sw[r] of the product in
    its epilogue.  Real weights only (the bulk case); complex boundary factors take the explicit elementwise pass, as
    does any ``matmul`` without the fused option (the CPU engines of the gloo tests).  ``left_t`` may be a row block
    (sharded theta build): ``sw`` is then the matching slice."""
    if sw.dim() == 2:
        sw = torch.diagonal(sw, 0)
    l, s, m = left_t.shape
    _m, t, r = right_t.shape
    if fused and not sw.is_complex():
        if direction == "right":
            return matmul(left_t.reshape(l * s, m), right_t.reshape(m, t * r), row_scale=(sw, s))
        return matmul(left_t.reshape(l * s, m), right_t.reshape(m, t * r), col_scale=(sw, r))
    if direction == "right":
        return matmul(torch.mul(left_t, sw.to(left_t.dtype).view(-1, 1, 1)).reshape(l * s, m), right_t.reshape(m, t * r))
    return matmul(left_t.reshape(l * s, m), torch.mul(right_t, sw.to(right_t.dtype).view(1, 1, -1)).reshape(m, t * r))


class _BaseContractModule(nn.Module):
    """dmrg.py:127-162: validates and stores the backend policy."""

    def __init__(self, contract_backend: str = "auto"):
        super().__init__()
        if contract_backend not in _BACKENDS:
            raise ValueError("contract backend must be 'auto'|'einsum'|'ncon'")
        self.contract_backend = contract_backend

    def _resolved(self, ref: torch.Tensor) -> str:
        return _CONTRACT_BACKEND


def _dev_tensor(t: torch.Tensor) -> torch.Tensor:
    return _lib._resolved(t.detach() if t.requires_grad else t)


def _same_dtype(where: str, *tensors: torch.Tensor) -> None:
    """The kernels read every operand with the dtype of the first: mixed dtypes are an error, not a cast."""
    dt = tensors[0].dtype
    if dt not in (torch.float64, torch.complex128):
        raise ValueError(f"[{where}] unsupported dtype {dt}: the contraction kernels take float64 or complex128")
    for t in tensors[1:]:
        if t.dtype != dt:
            raise ValueError(f"[{where}] dtype mismatch: {dt} vs {t.dtype}")


class ApplyMPO(_BaseContractModule):
    r"""H_eff . theta = L . M1 . M2 . R . theta for the two-site block (dmrg.py:165-248).

    ``psi_flat`` is the row-major flattening of theta[a, s1, s2, b]; ``L[w, a', a]``, ``M[w_l, w_r, s', s]``,
    ``R[w, b', b]`` (dmrg.py:188-207).  L and R may have any strides.
    """

    def forward(self, psi_flat: torch.Tensor, L: torch.Tensor, M1: torch.Tensor, M2: torch.Tensor,
                R: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _same_dtype("ApplyMPO", psi_flat, L, M1, M2, R)
        if psi_flat.device.type == "cpu" and L.device.type == "cuda":
            return self._forward_host(psi_flat, L, M1, M2, R, out)
        psi, L, M1, M2, R = (_dev_tensor(t) for t in (psi_flat, L, M1, M2, R))
        if not psi.is_contiguous():
            psi = psi.contiguous()
        if not M1.is_contiguous():
            M1 = M1.contiguous()
        if not M2.is_contiguous():
            M2 = M2.contiguous()
        n_in = L.shape[2] * M1.shape[3] * M2.shape[3] * R.shape[2]
        if psi.numel() != n_in:
            raise AssertionError(f"[ApplyMPO] psi has {psi.numel()} elements, the operator expects {n_in}")
        n_out = L.shape[1] * M1.shape[2] * M2.shape[2] * R.shape[1]
        if out is not None:
            if out.numel() != n_out or out.device != psi.device or out.dtype != psi.dtype \
                    or not out.is_contiguous():
                raise ValueError("[ApplyMPO] out shape/device/dtype mismatch.")      # dmrg.py:224-225
            res = out.view(-1)
        else:
            res = torch.empty(n_out, device=psi.device, dtype=psi.dtype)
        (li, lrow0), (ri, _r0) = _lib.identity_of(L), _lib.identity_of(R)
        if li >= 0 or ri >= 0:         # verified identity channels: their chi^3 GEMM slices are skipped
            _lib.ops.heff_apply_ident(psi.view(-1), L, M1, M2, R, res, li, ri, lrow0)
        else:
            _lib.ops.heff_apply(psi.view(-1), L, M1, M2, R, res)
        return res


    # ---- host-resident vector, device-resident operator ---------------------------------------------------
    _host_pipes: dict = {}

    def _forward_host(self, psi_host: torch.Tensor, L, M1, M2, R, out_host: Optional[torch.Tensor]) -> torch.Tensor:
        """``psi_flat`` (and ``out``) in PINNED HOST memory, L / M / R on the GPU: the upload of theta, the matvec
        and the download of the result are pipelined instead of serialised.  theta, viewed as a (chi_l, d1 d2 chi_r)
        matrix, goes up in COLUMN blocks on a copy stream (qsb_upload_columns), each followed by a stream-ordered
        flag write; an output tile of the step-1 GEMM reads exactly one column block and is gated on its flag, so
        compute starts after 1/8 of the upload (row-block / k-chunk gating, the mechanism of the multi-GPU fused
        exchange, is the fallback when the column blocks do not align with the tiles).  The matvec runs in row
        blocks of the left bond and each finished block is downloaded on a second copy stream while the next one
        computes.  Compute stays on the GPU -- this is
        not a CPU fallback; unpinned host tensors are rejected."""
        (li, lrow0), (ri, _r0) = _lib.identity_of(L), _lib.identity_of(R)
        L, M1, M2, R = (_dev_tensor(t) for t in (L, M1, M2, R))
        M1, M2 = M1.contiguous(), M2.contiguous()
        dev = L.device
        psi_host = psi_host.reshape(-1)
        if not psi_host.is_pinned() or not psi_host.is_contiguous():
            raise ValueError("[ApplyMPO] a host-resident psi_flat must be contiguous pinned memory")
        dt = psi_host.dtype
        chi_lo, chi_li = L.shape[1], L.shape[2]
        n_in = chi_li * M1.shape[3] * M2.shape[3] * R.shape[2]
        per_out = M1.shape[2] * M2.shape[2] * R.shape[1]
        n_out = chi_lo * per_out
        if psi_host.numel() != n_in:
            raise AssertionError(f"[ApplyMPO] psi has {psi_host.numel()} elements, the operator expects {n_in}")
        if out_host is None:
            out_host = torch.empty(n_out, dtype=dt).pin_memory()
        if out_host.numel() != n_out or out_host.dtype != dt or out_host.device.type != "cpu" \
                or not out_host.is_pinned() or not out_host.is_contiguous():
            raise ValueError("[ApplyMPO] out shape/device/dtype mismatch.")
        pipe_key = (dev, torch.cuda.current_stream(dev).cuda_stream)     # one pipeline per (device, launching stream)
        pipe = ApplyMPO._host_pipes.get(pipe_key)
        if pipe is None:
            pipe = {"flags": torch.zeros(64, dtype=torch.int32, device=dev), "epoch": 0,
                    "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev), "theta": None,
                    "out": None}
            ApplyMPO._host_pipes[pipe_key] = pipe
        for key, numel in (("theta", n_in), ("out", n_out)):
            buf = pipe[key]
            if buf is None or buf.numel() < numel or buf.dtype != dt:
                pipe[key] = torch.empty(numel, dtype=dt, device=dev)
        theta_dev, out_dev = pipe["theta"][:n_in], pipe["out"][:n_out]
        bk, bn = (8, 64) if dt.is_complex else (16, 128)
        N1 = n_in // chi_li                               # theta as a (chi_l, d1 d2 chi_r) matrix
        # preferred: COLUMN blocks of theta -- an output tile of step 1 needs one block only, so the first tiles start
        # after 1/8 of the upload; fallback: ROW blocks (k chunks), then a plain upload
        col_chunks = 16
        while col_chunks > 1 and (N1 % col_chunks != 0 or (N1 // col_chunks) % bn != 0):
            col_chunks //= 2
        row_chunks = 8
        while row_chunks > 1 and (chi_li % (bk * row_chunks) != 0):
            row_chunks //= 2
        mode = "col" if col_chunks > 1 else ("row" if row_chunks > 1 else "plain")
        bounds = _lib.row_block_bounds(chi_lo)            # row blocks of the result, big first, small last
        main = torch.cuda.current_stream(dev)
        es = psi_host.element_size()
        with torch.cuda.device(dev):
            pipe["h2d"].wait_stream(main)                  # the staging buffers are free once earlier work is done
            pipe["d2h"].wait_stream(main)
            pipe["epoch"] += 1
            epoch = pipe["epoch"]
            with torch.cuda.stream(pipe["h2d"]):
                if mode == "col":
                    _lib.upload_columns(psi_host, theta_dev, chi_li, N1, col_chunks, pipe["flags"], epoch)
                elif mode == "row":
                    chunk = n_in // row_chunks
                    for c in range(row_chunks):
                        _lib.push_shard_dma(psi_host[c * chunk:(c + 1) * chunk],
                                            [theta_dev.data_ptr() + c * chunk * es],
                                            [pipe["flags"].data_ptr() + 4 * c], epoch)
                else:
                    theta_dev.copy_(psi_host, non_blocking=True)
            if mode == "plain":
                main.wait_stream(pipe["h2d"])
            out_h = out_host.view(-1)
            for r_lo, r_hi in zip(bounds[:-1], bounds[1:]):
                Lb = L[:, r_lo:r_hi, :]
                ob = out_dev[r_lo * per_out:r_hi * per_out]
                row0 = lrow0 + r_lo                        # an identity channel of L, seen from this row block
                if mode == "col":
                    try:
                        # (the identity shortcut stays valid under gating: the step-1 GEMM of the other channels reads
                        # every block of theta, so all of it has landed before the mix reads theta in place)
                        _lib.ops.heff_apply_colgated(theta_dev, Lb, M1, M2, R, ob, pipe["flags"], col_chunks, epoch,
                                                     li, ri, row0)
                    except ValueError:                     # operands the TMA-staged (gated) GEMM cannot describe
                        main.wait_stream(pipe["h2d"])
                        mode = "plain"
                        _lib.ops.heff_apply_ident(theta_dev, Lb, M1, M2, R, ob, li, ri, row0)
                elif mode == "row":
                    try:
                        _lib.ops.heff_apply_gated(theta_dev, Lb, M1, M2, R, ob, pipe["flags"], row_chunks, 0, epoch, 0,
                                                  li, ri, row0)
                    except ValueError:
                        main.wait_stream(pipe["h2d"])
                        mode = "plain"
                        _lib.ops.heff_apply_ident(theta_dev, Lb, M1, M2, R, ob, li, ri, row0)
                else:
                    _lib.ops.heff_apply_ident(theta_dev, Lb, M1, M2, R, ob, li, ri, row0)
                done = torch.cuda.Event()
                done.record(main)
                pipe["d2h"].wait_event(done)
                with torch.cuda.stream(pipe["d2h"]):
                    out_h[r_lo * per_out:r_hi * per_out].copy_(ob, non_blocking=True)
            pipe["last_mode"] = mode
            main.wait_stream(pipe["d2h"])                  # the result is on the host once `main` gets here
            main.wait_stream(pipe["h2d"])
            pipe["done"] = torch.cuda.Event()
            pipe["done"].record(main)
        # CONTRACT: the copies are asynchronous.  `out_host` holds the result -- and `psi_host` may be
        # overwritten -- only after the launching stream has reached this point: synchronise the stream (or the device),
        # or call ApplyMPO.host_result_ready(device).synchronize().
        return out_host

    @classmethod
    def host_result_ready(cls, device=None) -> "torch.cuda.Event":
        """Event recorded after the last download of the most recent host-resident call on the current stream."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        pipe = cls._host_pipes.get((dev, torch.cuda.current_stream(dev).cuda_stream))
        if pipe is None or "done" not in pipe:
            raise RuntimeError("no host-resident ApplyMPO call has been issued on this (device, stream)")
        return pipe["done"]

    @classmethod
    def release_host_pipes(cls) -> None:
        """Free the device staging buffers (2 n elements per (device, stream)) of the host-resident path."""
        torch.cuda.synchronize()
        cls._host_pipes.clear()


class LeftEnvUpdate(_BaseContractModule):
    """L[p+1][w', c'_bra, c_ket] = sum L[p] . conj(A) . M . A  (dmrg.py:251-278); returns contiguous (W', chi', chi')."""

    def forward(self, Lp: torch.Tensor, M: torch.Tensor, Ap: torch.Tensor) -> torch.Tensor:
        _same_dtype("LeftEnvUpdate", Lp, M, Ap)
        Lp, M, Ap = (_dev_tensor(t) for t in (Lp, M, Ap))
        if not M.is_contiguous():
            M = M.contiguous()
        ident = _lib.identity_of(Lp)[0]
        out = torch.empty(M.shape[1], Ap.shape[2], Ap.shape[2], device=Lp.device, dtype=Lp.dtype)
        _lib.ops.env_left(Lp, M, Ap, out, ident)
        return _lib.tag_identity(out)


class RightEnvUpdate(_BaseContractModule):
    """R[p][w, a'_bra, a_ket] = sum M . R[p+1] . B . conj(B)  (dmrg.py:281-308); returns contiguous (W, chi, chi)."""

    def forward(self, M: torch.Tensor, Rnext: torch.Tensor, Bp1: torch.Tensor) -> torch.Tensor:
        _same_dtype("RightEnvUpdate", M, Rnext, Bp1)
        M, Rnext, Bp1 = (_dev_tensor(t) for t in (M, Rnext, Bp1))
        if not M.is_contiguous():
            M = M.contiguous()
        ident = _lib.identity_of(Rnext)[0]
        out = torch.empty(M.shape[0], Bp1.shape[0], Bp1.shape[0], device=Rnext.device, dtype=Rnext.dtype)
        _lib.ops.env_right(Rnext, M, Bp1, out, ident)
        return _lib.tag_identity(out)


class EnvironmentManager:
    """Caches L[0..N] and R[0..N]; boundaries are ones(W,1,1)  (dmrg.py:311-418)."""

    def __init__(self, mpo: List[torch.Tensor], device: Any, dtype: torch.dtype, contract_backend: str = "auto"):
        self.mpo = mpo
        self.Nsites = len(mpo)
        self.device = device
        self.dtype = dtype
        self.L_cache: List[Optional[torch.Tensor]] = [None] * (self.Nsites + 1)
        self.R_cache: List[Optional[torch.Tensor]] = [None] * (self.Nsites + 1)
        self.L_cache[0] = torch.ones(self.mpo[0].shape[0], 1, 1, device=device, dtype=dtype)
        self.R_cache[self.Nsites] = torch.ones(self.mpo[self.Nsites - 1].shape[1], 1, 1, device=device, dtype=dtype)
        self._left_upd = LeftEnvUpdate(contract_backend)
        self._right_upd = RightEnvUpdate(contract_backend)

    def get_L(self, site_idx: int) -> torch.Tensor:
        if self.L_cache[site_idx] is None:
            raise RuntimeError(f"L[{site_idx}] requested but not yet computed.")
        return self.L_cache[site_idx]

    def get_R(self, site_idx: int) -> torch.Tensor:
        if self.R_cache[site_idx] is None:
            raise RuntimeError(f"R[{site_idx}] requested but not yet computed.")
        return self.R_cache[site_idx]

    def update_L(self, site_idx: int, A_tensor: torch.Tensor) -> None:
        self.L_cache[site_idx + 1] = self._left_upd(self.get_L(site_idx), self.mpo[site_idx], A_tensor)

    def update_R(self, site_idx: int, B_tensor: torch.Tensor) -> None:
        self.R_cache[site_idx] = self._right_upd(self.mpo[site_idx], self.get_R(site_idx + 1), B_tensor)


class Sweeps:
    """Per-sweep schedule (dmrg.py:424-503)."""

    def __init__(self, numsweeps: int):
        if numsweeps <= 0:
            raise ValueError("Number of sweeps must be positive.")
        self.numsweeps = numsweeps
        self.maxdim = [10] * numsweeps
        self.noise = [0.0] * numsweeps
        self.krylov_dim = [4] * numsweeps
        self.cutoff = [1e-12] * numsweeps
        self.reortho = [False] * numsweeps
        self.maxreortho = [1] * numsweeps

    def set_schedule(self, **kwargs) -> None:
        for key, val in kwargs.items():
            if not hasattr(self, key):
                raise AttributeError(f"Sweeps object does not have parameter '{key}'")
            if len(val) != self.numsweeps:
                raise ValueError(f"Length of schedule for '{key}' ({len(val)}) "
                                 f"does not match the number of sweeps ({self.numsweeps})")
            setattr(self, key, val)

    def __repr__(self) -> str:
        rows = [f"Sweeps({self.numsweeps}):"]
        for name in ("maxdim", "noise", "krylov_dim", "cutoff", "reortho", "maxreortho"):
            rows.append(f"  {name}: {getattr(self, name)}")
        return "\n".join(rows) + "\n"


class DMRG(nn.Module):
    """Two-site DMRG driver (dmrg.py:506-802): same constructor, ``forward`` and ``history`` keys."""

    def __init__(self, mps: MPS, mpo: Any, *, device: Any = "cuda", dtype: torch.dtype = torch.float64,
                 contract_backend: str = "auto"):
        super().__init__()
        set_contract_backend(contract_backend, device)
        if torch.device(device).type != "cuda":
            raise RuntimeError("qsb200 DMRG runs on a CUDA device only (no CPU fallback); got device="
                               f"{device!r}")
        self.device = device
        self.dtype = dtype
        self.contract_backend = contract_backend
        self.mps = mps
        self.Nsites = len(mpo)
        self.Ms = _as_mpo_list(list(mpo), self.Nsites, device, dtype)
        self.ML = torch.ones(self.Ms[0].shape[0], 1, 1, device=device, dtype=dtype)
        self.MR = torch.ones(self.Ms[-1].shape[1], 1, 1, device=device, dtype=dtype)
        self.chid = self.Ms[0].shape[2]
        _check_mpo_shapes(self.Ms, self.ML, self.MR, self.chid)
        self.apply_mpo = ApplyMPO(contract_backend)
        self.history = self._empty_history()
        self.profile = False            # True: device-synchronised wall time per phase of every bond step -> self.timers
        self.timers = {"theta_build": 0.0, "lanczos": 0.0, "two_site_split": 0.0, "mps_update": 0.0, "env_update": 0.0}

    def _tick(self, phase: Optional[str], t0: float) -> float:
        """Profiling hook: with ``self.profile`` set, closes ``phase`` (started at ``t0``) and returns the new start."""
        if not self.profile:
            return 0.0
        import time
        torch.cuda.synchronize()
        now = time.perf_counter()
        if phase is not None:
            self.timers[phase] += now - t0
        return now

    @staticmethod
    def _empty_history() -> dict:
        return {"energies": [], "discarded_weights": [], "bond_dims": [], "sweeps": [], "directions": [],
                "sites": []}

    # -- one half sweep (dmrg.py:577-722) ---------------------------------------------------------
    def _sweep(self, direction: str, mps: MPS, env_manager: EnvironmentManager, sweep_params: dict,
               Ekeep: List[float]) -> List[float]:
        chi, noise, krydim = sweep_params["chi"], sweep_params["noise"], sweep_params["krydim"]
        cutoff, dispon, updateon = sweep_params["cutoff"], sweep_params["dispon"], sweep_params["updateon"]
        maxit, reortho, maxreortho = sweep_params["maxit"], sweep_params["reortho"], sweep_params["maxreortho"]
        sweep_idx, numsweeps = sweep_params["sweep_idx"], sweep_params["numsweeps"]
        A, B, sWeight = mps.A, mps.B, mps.sWeight
        d = self.chid
        sites = range(self.Nsites - 1) if direction == "right" else range(self.Nsites - 2, -1, -1)
        for p in sites:
            t0 = self._tick(None, 0.0)
            # 1. two-site wavefunction (dmrg.py:633-662)
            if direction == "right":
                left_t, right_t, sw = B[p], B[p + 1], sWeight[p]
            else:
                left_t, right_t, sw = A[p], A[p + 1], sWeight[p + 2]
            l, r = left_t.shape[0], right_t.shape[2]
            psi = build_theta(left_t, right_t, sw, direction)
            chil, chir = l, r
            psi_flat = psi.reshape(-1)
            t0 = self._tick("theta_build", t0)

            # 2. Lanczos ground state of H_eff (dmrg.py:664-675)
            if updateon:
                operator_args = (env_manager.get_L(p), self.Ms[p], self.Ms[p + 1], env_manager.get_R(p + 2))
                psi_flat, Entemp = lanczos_ground_state(
                    psi_flat, self.apply_mpo, operator_args, num_restarts=maxit, krylov_dim=krydim,
                    pro=reortho, pro_max_reorth=maxreortho)
                Ekeep.append(Entemp)
                self.history["energies"].append(float(Entemp))

            t0 = self._tick("lanczos", t0)

            # 3. optional noise, global RNG, uniform [0,1) like the reference (dmrg.py:677-682)
            if noise > 0:
                noise_vec = torch.rand_like(psi_flat)
                noise_vec.div_(torch.linalg.norm(noise_vec))
                psi_flat = psi_flat + noise_vec * noise
                psi_flat.div_(torch.linalg.norm(psi_flat))

            # 4. SVD + truncation + MPS update (dmrg.py:684-711); large blocks go through the Gram/eigh split (svd.py)
            Uk, S, Vhk, keep = two_site_split(psi_flat.view(chil * d, d * chir), chi, cutoff)
            t0 = self._tick("two_site_split", t0)
            self.history["discarded_weights"].append(torch.sum(S[keep:] ** 2).real.item())
            self.history["bond_dims"].append(int(keep))
            self.history["sweeps"].append(int(sweep_idx))
            self.history["directions"].append(direction)
            self.history["sites"].append(int(p))
            sv = S[:keep]
            denom = sv.norm().add_(torch.finfo(sv.dtype).eps)
            A[p] = nn.Parameter(Uk.reshape(chil, d, keep), requires_grad=False)
            sWeight[p + 1] = nn.Parameter(sv.div(denom), requires_grad=False)
            B[p + 1] = nn.Parameter(Vhk.reshape(keep, d, chir), requires_grad=False)

            t0 = self._tick("mps_update", t0)

            # 5. environment update (dmrg.py:713-717)
            if direction == "right":
                env_manager.update_L(p, A[p])
            else:
                env_manager.update_R(p + 1, B[p + 1])
            self._tick("env_update", t0)
            if dispon == 2 and updateon:
                print(f"Sweep: {sweep_idx} of {numsweeps}, Loc: {p}, Energy: {Ekeep[-1]:.6f}")
        return Ekeep

    # -- full schedule (dmrg.py:725-802) -----------------------------------------------------------
    def forward(self, sweeps: Sweeps, *, dispon=2, updateon=True, maxit=2):
        Ekeep: List[float] = []
        self.history = self._empty_history()
        N, d = self.Nsites, self.chid
        env_manager = EnvironmentManager(self.Ms, self.device, self.dtype, self.contract_backend)
        for p in range(N - 1, self.mps.ortho_center, -1):
            env_manager.update_R(p, self.mps.B[p])
        for k in range(1, sweeps.numsweeps + 1):
            params = {"chi": sweeps.maxdim[k - 1], "noise": sweeps.noise[k - 1], "krydim": sweeps.krylov_dim[k - 1],
                      "cutoff": sweeps.cutoff[k - 1], "reortho": sweeps.reortho[k - 1],
                      "maxreortho": sweeps.maxreortho[k - 1], "dispon": dispon, "updateon": updateon,
                      "maxit": maxit, "sweep_idx": k, "numsweeps": sweeps.numsweeps}
            Ekeep = self._sweep("right", self.mps, env_manager, params, Ekeep)
            self.mps.ortho_center = N - 1
            sw = self.mps.sWeight[N - 1]                                    # right boundary (dmrg.py:769-781)
            if sw.dim() == 2:
                sw = torch.diagonal(sw, 0)
            B_last = self.mps.B[N - 1]
            Atemp = (sw.to(B_last.dtype).view(-1, 1, 1) * B_last).reshape(B_last.shape[0] * d, B_last.shape[2])
            U, S, Vh = torch.linalg.svd(Atemp, full_matrices=False)
            eps = torch.finfo(S.dtype).eps
            self.mps.A[N - 1] = nn.Parameter(U.view(B_last.shape[0], d, B_last.shape[2]), requires_grad=False)
            self.mps.sWeight[N] = nn.Parameter((S.unsqueeze(1) * Vh).to(self.dtype) / (S.norm().add_(eps)),
                                               requires_grad=False)
            Ekeep = self._sweep("left", self.mps, env_manager, params, Ekeep)
            self.mps.ortho_center = 0
            sw = self.mps.sWeight[1]                                        # left boundary (dmrg.py:787-796)
            if sw.dim() == 2:
                sw = torch.diagonal(sw, 0)
            A0 = self.mps.A[0] * sw.to(self.mps.A[0].dtype).view(1, 1, -1)
            chil, _, chir = self.mps.A[0].shape
            U, S, Vh = torch.linalg.svd(A0.reshape(chil, d * chir), full_matrices=False)
            self.mps.B[0] = nn.Parameter(Vh.view(chil, d, chir), requires_grad=False)
            self.mps.sWeight[0] = nn.Parameter((U * S.unsqueeze(0)).to(self.dtype) / (S.norm().add_(eps)),
                                               requires_grad=False)
            if dispon == 1:
                print(f"Sweep: {k} of {sweeps.numsweeps}, Energy: {Ekeep[-1]:.12f}, "
                      f"Bond dim: {self.mps.A[N // 2].shape[-1]}, Noise: {sweeps.noise[k - 1]:.2e}, "
                      f"Cutoff: {sweeps.cutoff[k - 1]:.2e}")
        return Ekeep, self.mps
