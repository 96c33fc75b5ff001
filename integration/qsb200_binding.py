"""The binding a maintainer of ToelUl/quantum-simulation would add for the B200 kernels.

Self-contained: ctypes + torch only, no import of the ``quantum_simulation_b200`` package.  It mirrors the two
descriptor structs of ``include/qsb200.h`` FIELD FOR FIELD (and asserts their size against the library's own
``qsb_sizeof_*``), binds ``qsb_heff_apply`` / ``qsb_env_left`` / ``qsb_env_right``, and ``install()`` routes the
reference's own modules to them:

    quantum_simulation.dmrg.ApplyMPO.forward        (dmrg.py:178-248)  -> heff_apply
    quantum_simulation.dmrg.LeftEnvUpdate.forward   (dmrg.py:254-278)  -> env_left
    quantum_simulation.dmrg.RightEnvUpdate.forward  (dmrg.py:284-308)  -> env_right

CUDA tensors of dtype float64 / complex128 take the kernels; everything else keeps the reference's own code, so the
reference's ``DMRG``, ``EnvironmentManager`` and ``lanczos_ground_state`` (which calls
``linear_operator(v, *args, out=w)``, lanczos_solver.py:234-238) run unchanged on ``device='cuda'``.
tests/test_gpu_dropin_reference.py does exactly that against the golden energies of the stock reference.
"""
import ctypes as C
import os

import torch

_DEFAULT_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "quantum-simulation_b200", "lib",
                            "libqsb200.so")
_lib = None


class HeffDesc(C.Structure):                      # == qsb_heff_desc (include/qsb200.h)
    _fields_ = [("dtype", C.c_int),
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
                ("chunk_axis", C.c_int), ("chunk_local", C.c_int),
                ("L_identity", C.c_int), ("R_identity", C.c_int), ("L_row0", C.c_int),
                ("chunk_cols", C.c_int)]


class EnvDesc(C.Structure):                       # == qsb_env_desc
    _fields_ = [("dtype", C.c_int),
                ("W_in", C.c_int), ("W_out", C.c_int), ("d", C.c_int), ("chi_in", C.c_int), ("chi_out", C.c_int),
                ("env", C.c_void_p), ("env_stride", C.c_int64 * 3),
                ("M", C.c_void_p),
                ("site", C.c_void_p), ("site_stride", C.c_int64 * 3),
                ("out", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("chi_bra", C.c_int),
                ("site_bra", C.c_void_p), ("site_bra_stride", C.c_int64 * 3),
                ("env_identity", C.c_int)]


def load(path: str = None):
    """dlopen libqsb200.so once and check that the struct mirrors above match the library byte for byte."""
    global _lib
    if _lib is None:
        lib = C.CDLL(path or os.environ.get("QSB200_LIB", _DEFAULT_LIB))
        lib.qsb_sizeof_heff_desc.restype = C.c_size_t
        lib.qsb_sizeof_env_desc.restype = C.c_size_t
        lib.qsb_status_string.restype = C.c_char_p
        lib.qsb_env_identity_deviation.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                   C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p]
        assert lib.qsb_sizeof_heff_desc() == C.sizeof(HeffDesc), "HeffDesc mirror out of date"
        assert lib.qsb_sizeof_env_desc() == C.sizeof(EnvDesc), "EnvDesc mirror out of date"
        _lib = lib
    return _lib


def _check(rc, where):
    if rc == 0:
        return
    msg = f"[{where}] {load().qsb_status_string(int(rc)).decode()} (qsb_status {rc})"
    if rc == 1:
        raise AssertionError(msg)                 # bond mismatch: what the reference's einsum branch asserts
    raise ValueError(msg) if rc in (2, 3, 7) else RuntimeError(msg)   # never TypeError (lanczos_solver.py:236)


def _code(dt):
    if dt == torch.float64:
        return 0
    if dt == torch.complex128:
        return 1
    raise ValueError(f"qsb200: unsupported dtype {dt}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _i3(seq):
    return (C.c_int64 * 3)(*seq)


_ws = {}


def _workspace(dev, nbytes):
    ws = _ws.get(dev)
    if ws is None or ws.numel() < nbytes:
        _ws[dev] = None
        ws = _ws[dev] = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
    return ws


def _plain(t):
    t = t.detach() if t.requires_grad else t
    return t.resolve_conj() if t.is_conj() else t


# ---- identity channels (optional shortcut, see qsb_heff_desc.L_identity) ------------------------------------
IDENTITY_TOL = 1e-13


def _tag(env):
    """Record on the tensor which channel (if any) is the identity: one small kernel + a W-double read-back."""
    idx = -1
    if IDENTITY_TOL > 0 and env.shape[1] == env.shape[2]:
        dev_out = torch.empty(env.shape[0], dtype=torch.float64, device=env.device)
        with torch.cuda.device(env.device):
            _check(load().qsb_env_identity_deviation(_code(env.dtype), env.shape[0], env.shape[1], env.shape[2],
                                                     env.data_ptr(), _i3(env.stride()), 0, dev_out.data_ptr(),
                                                     _stream()), "identity_deviation")
        hit = torch.nonzero(dev_out.cpu() <= IDENTITY_TOL)
        idx = int(hit[0]) if hit.numel() else -1
    env._qsb_identity = (idx, _version(env))
    return env


def _version(t):
    try:
        return int(t._version)
    except RuntimeError:                           # inference tensors do not track versions
        return -1


def _identity_of(env):
    """The channel recorded by _tag, unless the tensor has been written in place since."""
    tag = getattr(env, "_qsb_identity", None)
    return tag[0] if tag is not None and tag[1] == _version(env) else -1


# ---- the three calls -----------------------------------------------------------------------------------------
def heff_apply(psi, L, M1, M2, R, out):
    lib = load()
    l_id, r_id = _identity_of(L), _identity_of(R)
    psi, L, M1, M2, R = (_plain(t) for t in (psi, L, M1, M2, R))
    psi, M1, M2 = psi.contiguous(), M1.contiguous(), M2.contiguous()
    d = HeffDesc()
    d.dtype = _code(psi.dtype)
    d.W_l, d.chi_l_out, d.chi_l_in = L.shape
    wl, d.W_m, d.d1_out, d.d1_in = M1.shape
    wm, d.W_r, d.d2_out, d.d2_in = M2.shape
    wr, d.chi_r_out, d.chi_r_in = R.shape
    assert wl == d.W_l and wm == d.W_m and wr == d.W_r          # dmrg.py:233-237
    d.theta, d.L, d.M1, d.M2, d.R, d.out = (t.data_ptr() for t in (psi, L, M1, M2, R, out))
    d.L_stride, d.R_stride = _i3(L.stride()), _i3(R.stride())
    d.L_identity, d.R_identity = l_id + 1, r_id + 1
    nb = C.c_size_t()
    _check(lib.qsb_heff_workspace_bytes(C.byref(d), C.byref(nb)), "ApplyMPO")
    ws = _workspace(psi.device, nb.value)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(psi.device):
        _check(lib.qsb_heff_apply(C.byref(d), _stream()), "ApplyMPO")
    return out


def _env_update(env, M, site, left):
    lib = load()
    ident = _identity_of(env)
    env, M, site = _plain(env), _plain(M).contiguous(), _plain(site)
    d = EnvDesc()
    d.dtype = _code(env.dtype)
    d.W_in = env.shape[0]
    if left:                                       # Lp (W, chi, chi), M (W, W', d, d), A (chi, d, chi')
        wi, d.W_out, d.d, _ = M.shape
        d.chi_in, _, d.chi_out = site.shape
    else:                                          # Rnext (W', chi', chi'), M (W, W', d, d), B (chi, d, chi')
        d.W_out, wi, d.d, _ = M.shape
        d.chi_out, _, d.chi_in = site.shape
    assert wi == d.W_in and env.shape[2] == d.chi_in
    out = torch.empty(d.W_out, d.chi_out, d.chi_out, dtype=env.dtype, device=env.device)
    d.env, d.M, d.site, d.out = env.data_ptr(), M.data_ptr(), site.data_ptr(), out.data_ptr()
    d.env_stride, d.site_stride = _i3(env.stride()), _i3(site.stride())
    d.env_identity = ident + 1
    nb = C.c_size_t()
    _check(lib.qsb_env_workspace_bytes(C.byref(d), C.byref(nb)), "EnvUpdate")
    ws = _workspace(env.device, nb.value)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(env.device):
        _check((lib.qsb_env_left if left else lib.qsb_env_right)(C.byref(d), _stream()), "EnvUpdate")
    return _tag(out)


def env_left(Lp, M, Ap):
    return _env_update(Lp, M, Ap, True)


def env_right(M, Rnext, Bp1):
    return _env_update(Rnext, M, Bp1, False)


def _ours(*tensors):
    return all(t.is_cuda and t.dtype in (torch.float64, torch.complex128) for t in tensors) and \
        len({t.dtype for t in tensors}) == 1


# ---- routing the reference's modules ---------------------------------------------------------------------------
def install(dmrg_module):
    """Patch ``quantum_simulation.dmrg`` in place; returns a function that undoes it.  The added branch is what
    INTEGRATION.md shows next to the 'ncon' / 'einsum' branches."""
    A, Lu, Ru = dmrg_module.ApplyMPO, dmrg_module.LeftEnvUpdate, dmrg_module.RightEnvUpdate
    saved = (A.forward, Lu.forward, Ru.forward)

    def apply_forward(self, psi_flat, L, M1, M2, R, out=None):
        if not _ours(psi_flat, L, M1, M2, R):
            return saved[0](self, psi_flat, L, M1, M2, R, out)
        n_out = L.shape[1] * M1.shape[2] * M2.shape[2] * R.shape[1]
        if out is None:
            out = psi_flat.new_empty(n_out)
        elif out.numel() != n_out or out.device != psi_flat.device or out.dtype != psi_flat.dtype \
                or not out.is_contiguous():
            raise ValueError("[ApplyMPO] out shape/device/dtype mismatch.")           # dmrg.py:224-225
        return heff_apply(psi_flat.reshape(-1), L, M1, M2, R, out.view(-1)).view(-1)

    def left_forward(self, Lp, M, Ap):
        return env_left(Lp, M, Ap) if _ours(Lp, M, Ap) else saved[1](self, Lp, M, Ap)

    def right_forward(self, M, Rnext, Bp1):
        return env_right(M, Rnext, Bp1) if _ours(M, Rnext, Bp1) else saved[2](self, M, Rnext, Bp1)

    A.forward, Lu.forward, Ru.forward = apply_forward, left_forward, right_forward

    def uninstall():
        A.forward, Lu.forward, Ru.forward = saved
    return uninstall
