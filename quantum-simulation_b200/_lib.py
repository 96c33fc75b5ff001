"""ctypes binding of the qsb200 C ABI (include/qsb200.h) and its PyTorch custom-op registration.

The shared library ``lib/libqsb200.so`` is built in-tree by ``build.py`` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing, or a tensor is not on a CUDA device, the call raises.

Every op is registered under ``torch.ops.qsb200`` with a CUDA-only implementation that unpacks
``Tensor -> (data_ptr, strides, current stream)`` and calls the ``extern "C"`` entry point; status codes are
mapped to the exception types the reference raises:

  QSB_ERR_SHAPE                    -> AssertionError   (dmrg.py:233-237)
  QSB_ERR_DTYPE / STRIDE / ARG     -> ValueError       (dmrg.py:224-225)
  QSB_ERR_DEVICE / WORKSPACE / LAUNCH -> RuntimeError
never TypeError (the reference's Lanczos swallows TypeError, lanczos_solver.py:236).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqsb200.so")

QSB_F64, QSB_C128, QSB_F32, QSB_C64 = 0, 1, 2, 3
_DTYPE_CODE = {torch.float64: QSB_F64, torch.complex128: QSB_C128, torch.float32: QSB_F32, torch.complex64: QSB_C64}

EXPORTS = [
    "qsb_version", "qsb_status_string", "qsb_last_cuda_error", "qsb_launch_counters",
    "qsb_heff_workspace_bytes", "qsb_heff_apply", "qsb_heff_flops", "qsb_heff_set_stage_events",
    "qsb_env_workspace_bytes", "qsb_env_left", "qsb_env_right", "qsb_env_flops",
    "qsb_lz_scratch_bytes", "qsb_lz_dot", "qsb_lz_axpy_norm", "qsb_lz_scale", "qsb_lz_combine",
    "qsb_strided_copy3", "qsb_gemm", "qsb_probe_dmma", "qsb_push_shard", "qsb_push_shard_dma", "qsb_upload_columns",
    "qsb_push_columns_dma",
    "qsb_heff_flops_executed", "qsb_env_identity_deviation",
    "qsb_sizeof_heff_desc", "qsb_sizeof_env_desc", "qsb_sizeof_gemm_desc",
]


class HeffDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("chi_l_out", C.c_int), ("chi_l_in", C.c_int), ("chi_r_out", C.c_int), ("chi_r_in", C.c_int),
        ("d1_out", C.c_int), ("d1_in", C.c_int), ("d2_out", C.c_int), ("d2_in", C.c_int),
        ("W_l", C.c_int), ("W_m", C.c_int), ("W_r", C.c_int),
        ("theta", C.c_void_p),
        ("L", C.c_void_p), ("L_stride", C.c_int64 * 3),
        ("M1", C.c_void_p), ("M2", C.c_void_p),
        ("R", C.c_void_p), ("R_stride", C.c_int64 * 3),
        ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("n_chunks", C.c_int), ("chunk_rot", C.c_int), ("chunk_epoch", C.c_int),
        ("chunk_flags", C.c_void_p),
        ("chunk_axis", C.c_int),
        ("chunk_local", C.c_int),
        ("L_identity", C.c_int), ("R_identity", C.c_int), ("L_row0", C.c_int), ("chunk_cols", C.c_int),
    ]


class EnvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("W_in", C.c_int), ("W_out", C.c_int), ("d", C.c_int), ("chi_in", C.c_int), ("chi_out", C.c_int),
        ("env", C.c_void_p), ("env_stride", C.c_int64 * 3),
        ("M", C.c_void_p),
        ("site", C.c_void_p), ("site_stride", C.c_int64 * 3),
        ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("chi_bra", C.c_int),
        ("site_bra", C.c_void_p), ("site_bra_stride", C.c_int64 * 3),
        ("env_identity", C.c_int),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("P", C.c_int), ("Z", C.c_int),
        ("A", C.c_void_p), ("a_ms", C.c_int64), ("a_ks", C.c_int64), ("a_ps", C.c_int64), ("a_zs", C.c_int64),
        ("conjA", C.c_int),
        ("B", C.c_void_p), ("b_ns", C.c_int64), ("b_ks", C.c_int64), ("b_ps", C.c_int64), ("b_zs", C.c_int64),
        ("conjB", C.c_int),
        ("C", C.c_void_p), ("c_ms", C.c_int64), ("c_ns", C.c_int64), ("c_zs", C.c_int64),
        ("force_path", C.c_int),
        ("accumulate", C.c_int),
        ("row_scale", C.c_void_p), ("row_div", C.c_int),
        ("col_scale", C.c_void_p), ("col_mod", C.c_int),
    ]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libqsb200.so (once) and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"qsb200: CUDA library not found at {LIB_PATH}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback).")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.qsb_version.restype = i32
    lib.qsb_status_string.restype = C.c_char_p
    lib.qsb_status_string.argtypes = [i32]
    lib.qsb_last_cuda_error.restype = i32
    lib.qsb_launch_counters.restype = None
    lib.qsb_launch_counters.argtypes = [C.POINTER(C.c_uint64), i32]
    lib.qsb_heff_workspace_bytes.argtypes = [C.POINTER(HeffDesc), C.POINTER(C.c_size_t)]
    lib.qsb_heff_apply.argtypes = [C.POINTER(HeffDesc), vp]
    lib.qsb_heff_flops.argtypes = [C.POINTER(HeffDesc)]
    lib.qsb_heff_flops.restype = C.c_double
    lib.qsb_heff_set_stage_events.argtypes = [C.POINTER(vp), i32]
    lib.qsb_env_workspace_bytes.argtypes = [C.POINTER(EnvDesc), C.POINTER(C.c_size_t)]
    lib.qsb_env_left.argtypes = [C.POINTER(EnvDesc), vp]
    lib.qsb_env_right.argtypes = [C.POINTER(EnvDesc), vp]
    lib.qsb_env_flops.argtypes = [C.POINTER(EnvDesc)]
    lib.qsb_env_flops.restype = C.c_double
    lib.qsb_lz_scratch_bytes.argtypes = [i32]
    lib.qsb_lz_scratch_bytes.restype = C.c_size_t
    lib.qsb_lz_dot.argtypes = [i32, i64, i32, vp, i64, vp, vp, i32, i32, vp, vp]
    lib.qsb_lz_axpy_norm.argtypes = [i32, i64, i32, vp, i64, vp, vp, i32, i32, i32, vp, vp]
    lib.qsb_lz_scale.argtypes = [i32, i64, vp, vp, vp, i32, vp]
    lib.qsb_lz_combine.argtypes = [i32, i64, i32, vp, i64, C.POINTER(C.c_double), vp, vp, i32, i32, vp, vp]
    lib.qsb_strided_copy3.argtypes = [i32, C.POINTER(i64), vp, C.POINTER(i64), vp, i32, vp]
    lib.qsb_gemm.argtypes = [C.POINTER(GemmDesc), vp]
    lib.qsb_push_shard.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(vp), i32, i32, vp, vp]
    lib.qsb_upload_columns.argtypes = [vp, vp, i64, i64, i32, i32, vp, i32, vp]
    lib.qsb_push_shard_dma.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(vp), i32, i32, vp]
    lib.qsb_push_columns_dma.argtypes = [vp, i64, i64, i32, i32, vp, C.POINTER(vp), C.POINTER(vp), i32, i32, vp]
    lib.qsb_probe_dmma.argtypes = [i32, C.POINTER(C.c_double), vp, vp]
    lib.qsb_heff_flops_executed.argtypes = [C.POINTER(HeffDesc)]
    lib.qsb_heff_flops_executed.restype = C.c_double
    lib.qsb_env_identity_deviation.argtypes = [i32, i32, i32, i32, vp, C.POINTER(i64), i32, vp, vp]
    for fn, st in (("qsb_sizeof_heff_desc", HeffDesc), ("qsb_sizeof_env_desc", EnvDesc),
                   ("qsb_sizeof_gemm_desc", GemmDesc)):
        getattr(lib, fn).restype = C.c_size_t
        if getattr(lib, fn)() != C.sizeof(st):          # a stale .so or a drifted mirror must not run
            raise RuntimeError(f"qsb200: {fn}() = {getattr(lib, fn)()} but the ctypes mirror has {C.sizeof(st)} bytes; "
                               "rebuild lib/libqsb200.so")
    for name in EXPORTS:          # every symbol include/qsb200.h declares must be exported
        getattr(lib, name)
    _lib = lib
    return lib


class nvtx_range:
    """NVTX range around a hot-path call (visible in nsys / ncu --nvtx), enabled with QSB200_NVTX=1: zero cost
    otherwise.  Ranges: qsb.heff_apply, qsb.env_left / env_right, qsb.lanczos, qsb.two_site_split."""
    enabled = os.environ.get("QSB200_NVTX", "0") not in ("", "0")

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if nvtx_range.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if nvtx_range.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def row_block_bounds(rows: int, min_rows: int = 128, max_blocks: int = 5) -> list:
    """Row blocks of a host-resident matvec's result, big first, small last: 1/2, 1/4, 1/8, ... of the rows while a
    block stays >= ``min_rows`` (one GEMM tile row).  Every block's download overlaps the next block's compute, so only
    the LAST, smallest download is exposed, at the price of few extra launches.  Returns the boundaries [0, ..., rows]."""
    bounds = [0]
    rest = int(rows)
    while len(bounds) < max_blocks - 1 and rest % 2 == 0 and rest // 2 >= min_rows and rest > 2 * min_rows:
        bounds.append(bounds[-1] + rest // 2)
        rest -= rest // 2
    if rest % 2 == 0 and rest // 2 >= min_rows:
        bounds.append(bounds[-1] + rest // 2)
    bounds.append(int(rows))
    return bounds


def status_string(code: int) -> str:
    return load().qsb_status_string(int(code)).decode()


def check(code: int, where: str) -> None:
    """Map a qsb_status to the reference's exception types."""
    if code == 0:
        return
    msg = f"[{where}] {status_string(code)} (qsb_status {code})"
    if code == 1:
        raise AssertionError(msg)
    if code in (2, 3, 7):
        raise ValueError(msg)
    if code == 6:
        msg += f"; cudaError {load().qsb_last_cuda_error()}"
    raise RuntimeError(msg)


def launch_counters(reset: bool = False) -> Dict[str, int]:
    """Kernel launches issued by the library since load / last reset."""
    out = (C.c_uint64 * 3)()
    load().qsb_launch_counters(out, 1 if reset else 0)
    return {"tma_gemm": int(out[0]), "cpasync_gemm": int(out[1]), "other": int(out[2]),
            "total": int(out[0] + out[1] + out[2])}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise ValueError(f"qsb200: unsupported dtype {dt}") from None


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if t.device.type != "cuda":
            raise RuntimeError("qsb200 kernels run on CUDA tensors only (no CPU fallback); got device "
                               f"{t.device}")
        if t.device != dev:
            raise ValueError("qsb200: all tensors of one call must live on the same device")
    return dev


def _resolved(t: torch.Tensor) -> torch.Tensor:
    """Materialise lazy conj/neg bits (torch.conj views) so data_ptr is the value the kernels read."""
    if t.is_conj():
        t = t.resolve_conj()
    if t.is_neg():
        t = t.resolve_neg()
    return t


# ---------------------------------------------------------------------------------------------
# workspaces: one per (device, stream), grown on demand, reused by every call on that stream.  Calls on one stream are
# ordered, so they can share scratch; calls on different streams of a device get their own (re-entrant per
# (device, stream)).  The C library keeps its stream-K scratch by the same key.
# ---------------------------------------------------------------------------------------------
_workspaces: Dict[tuple, torch.Tensor] = {}


def workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    with torch.cuda.device(dev):
        key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream().cuda_stream)
        ws = _workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            _workspaces[key] = None  # release before growing (the allocator keeps the block for THIS stream's later use)
            ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
            _workspaces[key] = ws
    return ws


def release_workspaces() -> None:
    _workspaces.clear()


# ---------------------------------------------------------------------------------------------
# descriptors
# ---------------------------------------------------------------------------------------------
IDENTITY_TOL = 1e-13      # max |env[w] - 1| below which a channel counts as the identity; 0 disables the shortcut
_IDENT_ATTR = "_qsb_identity"


def identity_channel(env: torch.Tensor, row0: int = 0, tol: Optional[float] = None) -> int:
    """Index of the first channel w with env[w] == identity (rows shifted by ``row0`` for a row shard) to within
    ``tol`` in the max norm, or -1.  One small kernel + one host read-back of W doubles."""
    tol = IDENTITY_TOL if tol is None else tol
    if tol <= 0 or env.dim() != 3 or env.shape[1] > env.shape[2] or env.dtype not in (torch.float64, torch.complex128):
        return -1
    dev = _require_cuda(env)
    env = _resolved(env)
    W, rows, cols = env.shape
    dev_out = torch.empty(W, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(load().qsb_env_identity_deviation(dtype_code(env.dtype), W, rows, cols, env.data_ptr(),
                                                (C.c_int64 * 3)(*env.stride()), int(row0), dev_out.data_ptr(),
                                                _stream()), "identity_channel")
    hit = torch.nonzero(dev_out.cpu() <= tol)
    return int(hit[0]) if hit.numel() else -1


def tag_identity(env: torch.Tensor, row0: int = 0) -> torch.Tensor:
    """Test ``env`` for an identity channel and remember the answer ON the tensor object (the reference's call
    signature ``ApplyMPO.forward(psi, L, M1, M2, R)`` has no room for it); returns ``env``."""
    setattr(env, _IDENT_ATTR, (identity_channel(env, row0), int(row0), _version_of(env)))
    return env


def _version_of(t: torch.Tensor) -> int:
    try:
        return int(t._version)
    except RuntimeError:            # inference tensors do not track versions
        return -1


def identity_of(env: torch.Tensor):
    """(channel or -1, row0) recorded by ``tag_identity``; untagged tensors, and tensors written in place since
    they were tagged, have no shortcut."""
    tag = getattr(env, _IDENT_ATTR, None)
    if tag is None or tag[2] != _version_of(env):
        return (-1, 0)
    return tag[:2]


def heff_desc(psi: torch.Tensor, L: torch.Tensor, M1: torch.Tensor, M2: torch.Tensor, R: torch.Tensor,
              out: Optional[torch.Tensor], ident=(-1, -1, 0)) -> HeffDesc:
    if L.dim() != 3 or R.dim() != 3 or M1.dim() != 4 or M2.dim() != 4:
        raise AssertionError("[ApplyMPO] expected L,R rank 3 and M1,M2 rank 4")
    d = HeffDesc()
    d.dtype = dtype_code(psi.dtype)
    d.W_l, d.chi_l_out, d.chi_l_in = L.shape
    Wl1, d.W_m, d.d1_out, d.d1_in = M1.shape
    Wm2, d.W_r, d.d2_out, d.d2_in = M2.shape
    Wr3, d.chi_r_out, d.chi_r_in = R.shape
    # the reference's einsum branch asserts exactly these (dmrg.py:233-237)
    assert Wl1 == d.W_l, "Mismatch in left MPO bond dimension"
    assert Wm2 == d.W_m, "Mismatch in middle MPO bond dimension"
    assert Wr3 == d.W_r, "Mismatch in right MPO bond dimension"
    d.theta = psi.data_ptr()
    d.L = L.data_ptr()
    d.L_stride = (C.c_int64 * 3)(*L.stride())
    d.M1 = M1.data_ptr()
    d.M2 = M2.data_ptr()
    d.R = R.data_ptr()
    d.R_stride = (C.c_int64 * 3)(*R.stride())
    d.out = out.data_ptr() if out is not None else None
    d.L_identity, d.R_identity, d.L_row0 = ident[0] + 1, ident[1] + 1, ident[2]
    return d


def heff_flops(psi, L, M1, M2, R) -> float:
    return float(load().qsb_heff_flops(C.byref(heff_desc(psi, L, M1, M2, R, None))))


def heff_flops_executed(psi, L, M1, M2, R, ident=(-1, -1, 0)) -> float:
    """Flops the kernels execute for this matvec (identity channels skipped) -- the utilisation numerator."""
    return float(load().qsb_heff_flops_executed(C.byref(heff_desc(psi, L, M1, M2, R, None, ident))))


def _impl_heff_apply(psi, L, M1, M2, R, out, gate=None, ident=(-1, -1, 0)) -> None:
    lib = load()
    dev = _require_cuda(psi, L, M1, M2, R, out)
    d = heff_desc(psi, L, M1, M2, R, out, ident)
    if gate is not None:                   # (flags tensor, n_chunks, rot, epoch): fused all-gather -> matvec
        flags, d.n_chunks, d.chunk_rot, d.chunk_epoch, d.chunk_local = gate[:5]
        d.chunk_axis = gate[5] if len(gate) > 5 else 0
        d.chunk_cols = gate[6] if len(gate) > 6 else 0
        d.chunk_flags = flags.data_ptr()
    nb = C.c_size_t(0)
    check(lib.qsb_heff_workspace_bytes(C.byref(d), C.byref(nb)), "ApplyMPO")
    ws = workspace(dev, nb.value)
    d.workspace = ws.data_ptr()
    d.workspace_bytes = ws.numel()
    with torch.cuda.device(dev), nvtx_range("qsb.heff_apply"):
        check(lib.qsb_heff_apply(C.byref(d), _stream()), "ApplyMPO")


def _impl_heff_apply_ident(psi, L, M1, M2, R, out, l_ident: int, r_ident: int, l_row0: int) -> None:
    _impl_heff_apply(psi, L, M1, M2, R, out, ident=(l_ident, r_ident, l_row0))


def _impl_heff_apply_gated(psi, L, M1, M2, R, out, flags, n_chunks: int, rot: int, epoch: int, local: int,
                           l_ident: int = -1, r_ident: int = -1, l_row0: int = 0) -> None:
    _impl_heff_apply(psi, L, M1, M2, R, out, gate=(flags, n_chunks, rot, epoch, local),
                     ident=(l_ident, r_ident, l_row0))


def _impl_heff_apply_colgated(psi, L, M1, M2, R, out, flags, n_chunks: int, epoch: int,
                              l_ident: int = -1, r_ident: int = -1, l_row0: int = 0) -> None:
    _impl_heff_apply(psi, L, M1, M2, R, out, gate=(flags, n_chunks, 0, epoch, 0, 1),
                     ident=(l_ident, r_ident, l_row0))


def _impl_heff_apply_gridgated(psi, L, M1, M2, R, out, flags, n_rows: int, n_cols: int, rot: int, epoch: int,
                               l_ident: int = -1, r_ident: int = -1, l_row0: int = 0) -> None:
    """theta arrives in n_rows row blocks (source ranks) x n_cols column blocks, flags[r * n_cols + c] >= epoch."""
    _impl_heff_apply(psi, L, M1, M2, R, out, gate=(flags, n_rows, rot, epoch, 0, 2, n_cols),
                     ident=(l_ident, r_ident, l_row0))


def push_columns_dma(src: torch.Tensor, rows: int, row_elems: int, n_chunks: int, ready_flags: torch.Tensor,
                     dst_ptrs: Sequence[int], flag_ptrs: Sequence[int], epoch: int) -> None:
    """Forward a device block to peers column block by column block as the blocks land (see qsb_push_columns_dma), on the
    CURRENT stream (callers use side streams)."""
    dev = _require_cuda(src, ready_flags)
    cnt = len(dst_ptrs)
    dp = (C.c_void_p * max(cnt, 1))(*[C.c_void_p(int(p)) for p in dst_ptrs])
    fp = (C.c_void_p * max(cnt, 1))(*[C.c_void_p(int(p)) for p in flag_ptrs])
    with torch.cuda.device(dev):
        check(load().qsb_push_columns_dma(src.data_ptr(), rows, row_elems, src.element_size(), n_chunks,
                                          ready_flags.data_ptr(), dp, fp, cnt, epoch, _stream()), "push_columns_dma")


def upload_columns(src_host: torch.Tensor, dst: torch.Tensor, rows: int, row_elems: int, n_chunks: int,
                   flags: torch.Tensor, epoch: int) -> None:
    """Pinned host (rows x row_elems) matrix -> device in column blocks + per-block flags, on the CURRENT stream."""
    if src_host.device.type != "cpu" or not src_host.is_pinned():
        raise ValueError("upload_columns: the source must be pinned host memory")
    dev = _require_cuda(dst, flags)
    with torch.cuda.device(dev):
        check(load().qsb_upload_columns(src_host.data_ptr(), dst.data_ptr(), rows, row_elems, src_host.element_size(),
                                        n_chunks, flags.data_ptr(), epoch, _stream()), "upload_columns")


def _impl_push_shard(src, dst_ptrs: Sequence[int], flag_ptrs: Sequence[int], epoch: int, ticket) -> None:
    dev = _require_cuda(src, ticket)
    world = len(dst_ptrs)
    dp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in dst_ptrs])
    fp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in flag_ptrs])
    with torch.cuda.device(dev):
        check(load().qsb_push_shard(src.data_ptr(), src.numel() * src.element_size(), dp, fp, world, epoch,
                                    ticket.data_ptr(), _stream()), "push_shard")


def push_shard_dma(src: torch.Tensor, dst_ptrs: Sequence[int], flag_ptrs: Sequence[int], epoch: int) -> None:
    """Copy-engine push, destinations in the given order, on the CURRENT stream (callers use side streams).
    ``src`` is a CUDA tensor or a pinned host tensor."""
    if src.device.type == "cuda":
        dev = src.device
    else:
        if not src.is_pinned():
            raise ValueError("push_shard_dma: a host source must be pinned memory")
        dev = torch.device("cuda", torch.cuda.current_device())
    world = len(dst_ptrs)
    dp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in dst_ptrs])
    fp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in flag_ptrs])
    with torch.cuda.device(dev):
        check(load().qsb_push_shard_dma(src.data_ptr(), src.numel() * src.element_size(), dp, fp, world, epoch,
                                        _stream()), "push_shard_dma")


def env_desc(env: torch.Tensor, M: torch.Tensor, site: torch.Tensor, out: Optional[torch.Tensor],
             left: bool, site_bra: Optional[torch.Tensor] = None, ident: int = -1) -> EnvDesc:
    if env.dim() != 3 or M.dim() != 4 or site.dim() != 3:
        raise AssertionError("[EnvUpdate] expected env rank 3, M rank 4, site tensor rank 3")
    d = EnvDesc()
    d.dtype = dtype_code(env.dtype)
    W_in, chi_a, chi_b = env.shape
    if left:
        Wi, Wo, d1, d2 = M.shape
        ci, ds, co = site.shape
    else:
        Wo, Wi, d1, d2 = M.shape
        co, ds, ci = site.shape
    assert Wi == W_in, "Mismatch between environment and MPO bond dimension"
    assert d1 == d2 == ds, "Mismatch in physical dimension"
    assert ci == chi_b, "Mismatch between environment and MPS bond dimension"
    if site_bra is None:
        assert chi_a == chi_b, "environment must be square in its two MPS bonds"
        d.chi_bra = 0
        d.site_bra = None
    else:                       # row-block form (multi-GPU sharding over the bra index)
        rows, ds2, cx = site_bra.shape
        assert ds2 == ds and cx == (co if left else ci), "site_bra does not match the site tensor"
        if left:
            assert chi_a == rows, "left row-block update: env rows must match site_bra rows"
        else:
            assert chi_a == chi_b, "right update needs the full (square) old environment"
        d.chi_bra = rows
        d.site_bra = site_bra.data_ptr()
        d.site_bra_stride = (C.c_int64 * 3)(*site_bra.stride())
    d.W_in, d.W_out, d.d, d.chi_in, d.chi_out = W_in, Wo, ds, ci, co
    d.env = env.data_ptr()
    d.env_stride = (C.c_int64 * 3)(*env.stride())
    d.M = M.data_ptr()
    d.site = site.data_ptr()
    d.site_stride = (C.c_int64 * 3)(*site.stride())
    d.out = out.data_ptr() if out is not None else None
    d.env_identity = ident + 1 if site_bra is None else 0
    return d


def env_flops(env, M, site, left: bool) -> float:
    return float(load().qsb_env_flops(C.byref(env_desc(env, M, site, None, left))))


def _impl_env(env, M, site, out, left: bool, site_bra=None, ident: int = -1) -> None:
    lib = load()
    dev = _require_cuda(env, M, site, out)
    d = env_desc(env, M, site, out, left, site_bra, ident)
    nb = C.c_size_t(0)
    check(lib.qsb_env_workspace_bytes(C.byref(d), C.byref(nb)), "EnvUpdate")
    ws = workspace(dev, nb.value)
    d.workspace = ws.data_ptr()
    d.workspace_bytes = ws.numel()
    with torch.cuda.device(dev), nvtx_range("qsb.env_left" if left else "qsb.env_right"):
        check((lib.qsb_env_left if left else lib.qsb_env_right)(C.byref(d), _stream()),
              "LeftEnvUpdate" if left else "RightEnvUpdate")


def _impl_env_left(env, M, site, out, ident: int = -1) -> None:
    _impl_env(env, M, site, out, True, ident=ident)


def _impl_env_right(env, M, site, out, ident: int = -1) -> None:
    _impl_env(env, M, site, out, False, ident=ident)


def _impl_env_left_rows(env, M, site, site_bra, out) -> None:
    _impl_env(env, M, site, out, True, site_bra)


def _impl_env_right_rows(env, M, site, site_bra, out) -> None:
    _impl_env(env, M, site, out, False, site_bra)


def _impl_lz_dot(X, m: int, ldx: int, w, n: int, scal, out_idx: int, flags: int, scratch) -> None:
    dev = _require_cuda(X, w, scal, scratch)
    with torch.cuda.device(dev):
        check(load().qsb_lz_dot(dtype_code(w.dtype), n, m, X.data_ptr(), ldx, w.data_ptr(), scal.data_ptr(),
                                out_idx, flags, scratch.data_ptr(), _stream()), "lanczos.dot")


def _impl_lz_axpy_norm(X, m: int, ldx: int, w, n: int, scal, coef_idx: int, norm_idx: int, flags: int,
                       scratch) -> None:
    dev = _require_cuda(X, w, scal, scratch)
    with torch.cuda.device(dev):
        check(load().qsb_lz_axpy_norm(dtype_code(w.dtype), n, m, X.data_ptr(), ldx, w.data_ptr(),
                                      scal.data_ptr(), coef_idx, norm_idx, flags, scratch.data_ptr(), _stream()),
              "lanczos.axpy_norm")


def _impl_lz_scale(src, dst, n: int, scal, idx: int) -> None:
    dev = _require_cuda(src, dst, scal)
    with torch.cuda.device(dev):
        check(load().qsb_lz_scale(dtype_code(src.dtype), n, src.data_ptr(), dst.data_ptr(), scal.data_ptr(), idx,
                                  _stream()), "lanczos.scale")


def _impl_lz_combine(X, m: int, ldx: int, y: Sequence[float], dst, n: int, scal, norm_idx: int, flags: int,
                     scratch) -> None:
    dev = _require_cuda(X, dst, scal, scratch)
    yy = (C.c_double * len(y))(*[float(v) for v in y])
    with torch.cuda.device(dev):
        check(load().qsb_lz_combine(dtype_code(dst.dtype), n, m, X.data_ptr(), ldx, yy, dst.data_ptr(),
                                    scal.data_ptr(), norm_idx, flags, scratch.data_ptr(), _stream()), "lanczos.combine")


def set_stage_events(events: Optional[Sequence[torch.cuda.Event]]) -> None:
    """Profiling hook of qsb_heff_apply: four recorded-once torch events, or None to clear."""
    lib = load()
    if not events:
        check(lib.qsb_heff_set_stage_events(None, 0), "set_stage_events")
        return
    arr = (C.c_void_p * 4)(*[C.c_void_p(e.cuda_event) for e in events])
    check(lib.qsb_heff_set_stage_events(arr, 4), "set_stage_events")


def lz_scratch_bytes() -> int:
    return int(load().qsb_lz_scratch_bytes(64))


# ---------------------------------------------------------------------------------------------
# plain GEMM (tests, peak probe)
# ---------------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, B: torch.Tensor, *, conjA: bool = False, conjB: bool = False,
         force_path: int = 0, out: Optional[torch.Tensor] = None, accumulate: bool = False,
         row_scale=None, col_scale=None) -> torch.Tensor:
    """C[z] = opA(A[z]) @ opB(B[z]) on the library's FP64 tensor-core core.  A: (Z, M, K) or (M, K);
    B: (Z, K, N) or (K, N); any strides with one unit stride per matrix.
    ``row_scale=(v, div)`` / ``col_scale=(v, mod)``: real float64 device vectors fused into the epilogue,
    C[m][n] *= v_r[m // div] * v_c[n % mod] (the Schmidt weights of the theta build)."""
    lib = load()
    dev = _require_cuda(A, B)
    squeeze = A.dim() == 2
    if squeeze:
        A, B = A.unsqueeze(0), B.unsqueeze(0)
    Z, M, K = A.shape
    Zb, Kb, N = B.shape
    assert Kb == K and Zb in (1, Z)
    if out is None:
        if accumulate:
            raise ValueError("gemm: accumulate=True needs `out`")
        out = torch.empty(Z, M, N, dtype=A.dtype, device=dev)
    else:
        out = out.unsqueeze(0) if squeeze else out
        if tuple(out.shape) != (Z, M, N) or not out.is_contiguous() or out.dtype != A.dtype:
            raise ValueError("gemm: `out` must be a contiguous (Z, M, N) tensor of the operands' dtype")
    g = GemmDesc()
    g.dtype = dtype_code(A.dtype)
    g.M, g.N, g.K, g.P, g.Z = M, N, K, 1, Z
    g.A, g.a_zs, g.a_ms, g.a_ks, g.a_ps, g.conjA = A.data_ptr(), A.stride(0), A.stride(1), A.stride(2), 0, int(conjA)
    g.B, g.b_zs, g.b_ks, g.b_ns, g.b_ps, g.conjB = (B.data_ptr(), B.stride(0) if Zb == Z else 0, B.stride(1),
                                                   B.stride(2), 0, int(conjB))
    g.C, g.c_zs, g.c_ms, g.c_ns = out.data_ptr(), M * N, N, 1
    g.force_path = force_path
    g.accumulate = int(accumulate)
    keep = []
    for name, sc, extent in (("row", row_scale, M), ("col", col_scale, N)):
        if sc is None:
            continue
        v, q = sc
        v = v.detach().to(device=dev, dtype=torch.float64).contiguous()
        q = int(q)
        need = (extent + q - 1) // q if name == "row" else min(q, extent)
        if q < 1 or v.numel() < need:
            raise ValueError(f"gemm: {name}_scale has {v.numel()} entries, {need} needed")
        keep.append(v)
        if name == "row":
            g.row_scale, g.row_div = v.data_ptr(), q
        else:
            g.col_scale, g.col_mod = v.data_ptr(), q
    with torch.cuda.device(dev):
        check(lib.qsb_gemm(C.byref(g), _stream()), "gemm")
    return out[0] if squeeze else out


def probe_dmma_tflops(iters: int = 20000, reps: int = 3) -> float:
    """Register-resident DMMA issue rate of this GPU in TFLOP/s (the FP64 tensor-pipe peak)."""
    lib = load()
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    flops = C.c_double(0.0)
    best = 0.0
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.qsb_probe_dmma(iters, C.byref(flops), sink.data_ptr(), _stream()), "probe_dmma")
        e1.record()
        e1.synchronize()
        if r > 0:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


# ---------------------------------------------------------------------------------------------
# torch custom-op registration (CUDA only)
# ---------------------------------------------------------------------------------------------
_torch_lib: Optional[torch.library.Library] = None


def register_ops() -> None:
    global _torch_lib
    if _torch_lib is not None:
        return
    tl = torch.library.Library("qsb200", "DEF")
    tl.define("heff_apply(Tensor psi, Tensor L, Tensor M1, Tensor M2, Tensor R, Tensor(a!) out) -> ()")
    tl.define("heff_apply_ident(Tensor psi, Tensor L, Tensor M1, Tensor M2, Tensor R, Tensor(a!) out, int l_ident, "
              "int r_ident, int l_row0) -> ()")
    tl.define("heff_apply_gated(Tensor psi, Tensor L, Tensor M1, Tensor M2, Tensor R, Tensor(a!) out, Tensor flags, "
              "int n_chunks, int rot, int epoch, int local, int l_ident=-1, int r_ident=-1, int l_row0=0) -> ()")
    tl.define("heff_apply_colgated(Tensor psi, Tensor L, Tensor M1, Tensor M2, Tensor R, Tensor(a!) out, Tensor flags, "
              "int n_chunks, int epoch, int l_ident=-1, int r_ident=-1, int l_row0=0) -> ()")
    tl.define("heff_apply_gridgated(Tensor psi, Tensor L, Tensor M1, Tensor M2, Tensor R, Tensor(a!) out, Tensor flags, "
              "int n_rows, int n_cols, int rot, int epoch, int l_ident=-1, int r_ident=-1, int l_row0=0) -> ()")
    tl.define("push_shard(Tensor src, int[] dst_ptrs, int[] flag_ptrs, int epoch, Tensor(a!) ticket) -> ()")
    tl.define("env_left(Tensor env, Tensor M, Tensor site, Tensor(a!) out, int ident=-1) -> ()")
    tl.define("env_right(Tensor env, Tensor M, Tensor site, Tensor(a!) out, int ident=-1) -> ()")
    tl.define("env_left_rows(Tensor env, Tensor M, Tensor site, Tensor site_bra, Tensor(a!) out) -> ()")
    tl.define("env_right_rows(Tensor env, Tensor M, Tensor site, Tensor site_bra, Tensor(a!) out) -> ()")
    tl.define("lz_dot(Tensor X, int m, int ldx, Tensor w, int n, Tensor(a!) scal, int out_idx, int flags, "
              "Tensor(b!) scratch) -> ()")
    tl.define("lz_axpy_norm(Tensor X, int m, int ldx, Tensor(a!) w, int n, Tensor(b!) scal, int coef_idx, "
              "int norm_idx, int flags, Tensor(c!) scratch) -> ()")
    tl.define("lz_scale(Tensor src, Tensor(a!) dst, int n, Tensor scal, int idx) -> ()")
    tl.define("lz_combine(Tensor X, int m, int ldx, float[] y, Tensor(a!) dst, int n, Tensor(b!) scal, "
              "int norm_idx, int flags, Tensor(c!) scratch) -> ()")
    tl.impl("heff_apply", _impl_heff_apply, "CUDA")
    tl.impl("heff_apply_ident", _impl_heff_apply_ident, "CUDA")
    tl.impl("heff_apply_gated", _impl_heff_apply_gated, "CUDA")
    tl.impl("heff_apply_colgated", _impl_heff_apply_colgated, "CUDA")
    tl.impl("heff_apply_gridgated", _impl_heff_apply_gridgated, "CUDA")
    tl.impl("push_shard", _impl_push_shard, "CUDA")
    tl.impl("env_left", _impl_env_left, "CUDA")
    tl.impl("env_right", _impl_env_right, "CUDA")
    tl.impl("env_left_rows", _impl_env_left_rows, "CUDA")
    tl.impl("env_right_rows", _impl_env_right_rows, "CUDA")
    tl.impl("lz_dot", _impl_lz_dot, "CUDA")
    tl.impl("lz_axpy_norm", _impl_lz_axpy_norm, "CUDA")
    tl.impl("lz_scale", _impl_lz_scale, "CUDA")
    tl.impl("lz_combine", _impl_lz_combine, "CUDA")
    _torch_lib = tl


register_ops()
ops = torch.ops.qsb200
