"""GPU parity of the contraction kernels (through the reference-facing API -> torch.ops -> C ABI) against the
golden vectors of the unmodified reference and against the oracle on seeded inputs.
Tolerance (BASELINE.json north_star): matvec relative error <= 1e-12 in fp64 / complex128."""
import pytest
import torch

from conftest import load_golden, golden_mpo, ncon_layout, rel_err
from oracle import dmrg_oracle as orc
import quantum_simulation_b200 as qsb
from quantum_simulation_b200 import _lib

pytestmark = pytest.mark.gpu
TOL = 1e-12
DEV = "cuda"


def _names(z):
    return [str(s) for s in z["names"]]


def randn(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, dtype=dtype, generator=g)


# ----------------------------------------------------------------------------- plain GEMM core
@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 7, 5, 3), (3, 130, 70, 33), (2, 128, 128, 64),
                                   (1, 257, 129, 100), (2, 64, 200, 17)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_core_all_layouts(dtype, shape, ta, tb):
    Z, M, N, K = shape
    A = randn((Z, M, K), dtype, 1)
    B = randn((Z, K, N), dtype, 2)
    ref = torch.matmul(A, B)
    Ad = (A.transpose(1, 2).contiguous().transpose(1, 2) if ta else A).to(DEV)
    Bd = (B.transpose(1, 2).contiguous().transpose(1, 2) if tb else B).to(DEV)
    out = _lib.gemm(Ad, Bd)
    assert rel_err(out, ref) <= TOL
    if dtype.is_complex:
        out = _lib.gemm(Ad, Bd, conjA=True, conjB=True)
        assert rel_err(out, torch.matmul(A.conj(), B.conj())) <= TOL


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(2, 256, 128, 64), (1, 144, 80, 100), (3, 128, 64, 8), (1, 16, 16, 4),
                                   (1, 400, 272, 130), (5, 512, 1024, 512)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_core_tma_path(dtype, shape, ta, tb):
    """Aligned operands must take the TMA-staged kernel (force_path=1 errors out if they cannot) and agree
    with the cp.async kernel (force_path=2) and with torch."""
    Z, M, N, K = shape
    A = randn((Z, M, K), dtype, 3)
    B = randn((Z, K, N), dtype, 4)
    ref = torch.matmul(A, B)
    Ad = (A.transpose(1, 2).contiguous().transpose(1, 2) if ta else A).to(DEV)
    Bd = (B.transpose(1, 2).contiguous().transpose(1, 2) if tb else B).to(DEV)
    before = qsb.launch_counters()["tma_gemm"]
    out = _lib.gemm(Ad, Bd, force_path=1)
    assert qsb.launch_counters()["tma_gemm"] == before + 1
    assert rel_err(out, ref) <= TOL
    out2 = _lib.gemm(Ad, Bd, force_path=2)
    assert rel_err(out2, ref) <= TOL
    if dtype.is_complex:
        out = _lib.gemm(Ad, Bd, conjA=True, conjB=False, force_path=1)
        assert rel_err(out, torch.matmul(A.conj(), B)) <= TOL


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(1, 1024, 1024, 4096), (1, 2048, 2048, 512), (1, 3072, 2560, 256),
                                   (5, 640, 1152, 320), (1, 128, 128, 8192), (2, 256, 384, 1040)])
def test_gemm_stream_k_schedules(dtype, shape):
    """Tile counts that are not a multiple of the SM count exercise the stream-K tail of the persistent schedule
    (contributor / finisher fix-up); the result must equal cuBLAS and be bit-reproducible run to run."""
    Z, M, N, K = shape
    if dtype.is_complex and M * N * K > 3e9:
        K //= 2
    g = torch.Generator(device=DEV).manual_seed(7)
    A = torch.randn(Z, M, K, dtype=dtype, device=DEV, generator=g)
    B = torch.randn(Z, K, N, dtype=dtype, device=DEV, generator=g)
    out = _lib.gemm(A, B, force_path=1)
    ref = torch.matmul(A, B)
    assert rel_err(out, ref) <= TOL
    out2 = _lib.gemm(A, B, force_path=1)
    assert torch.equal(out, out2)


def test_gemm_unaligned_falls_back_to_cpasync():
    A = randn((1, 33, 17), torch.float64, 5).to(DEV)
    B = randn((1, 17, 9), torch.float64, 6).to(DEV)
    before = qsb.launch_counters()
    out = _lib.gemm(A, B)
    after = qsb.launch_counters()
    assert after["cpasync_gemm"] == before["cpasync_gemm"] + 1 and after["tma_gemm"] == before["tma_gemm"]
    assert rel_err(out, torch.matmul(A, B)) <= TOL
    with pytest.raises(ValueError):
        _lib.gemm(A, B, force_path=1)


# ----------------------------------------------------------------------------- H_eff matvec
@pytest.mark.parametrize("name", _names(load_golden("matvec")))
def test_matvec_golden(name):
    z = load_golden("matvec")
    g = {k: torch.from_numpy(z[f"{name}/{k}"]) for k in ("psi", "L", "M1", "M2", "R", "out")}
    L, R = g["L"], g["R"]
    Ld, Rd = L.to(DEV), R.to(DEV)
    if str(z[f"{name}/layout"]) == "ncon":
        Ld, Rd = ncon_layout(Ld), ncon_layout(Rd)
        assert Ld.stride()[1] == 1 or Ld.shape[1] == 1
    op = qsb.ApplyMPO("ncon")
    out = op(g["psi"].to(DEV), Ld, g["M1"].to(DEV), g["M2"].to(DEV), Rd)
    assert out.shape == g["out"].shape
    assert rel_err(out, g["out"]) <= TOL
    buf = torch.empty_like(out)
    res = op(g["psi"].to(DEV), Ld, g["M1"].to(DEV), g["M2"].to(DEV), Rd, out=buf)
    assert res.data_ptr() == buf.data_ptr() and rel_err(buf, g["out"]) <= TOL
    with pytest.raises(ValueError, match="out shape/device/dtype mismatch"):
        op(g["psi"].to(DEV), Ld, g["M1"].to(DEV), g["M2"].to(DEV), Rd, out=torch.empty(3, device=DEV, dtype=buf.dtype))


@pytest.mark.parametrize("dtype,mk", [(torch.float64, "heis_f64_L100"), (torch.complex128, "heis_c128_L100"),
                                      (torch.complex128, "tfim_c128_L100_g0.5"), (torch.float64, "j1j2_f64_8x6"),
                                      (torch.float64, "hubbard_f64_L64"), (torch.complex128, "hubbard_f64_L64")])
@pytest.mark.parametrize("chi_l,chi_r,layout", [(96, 96, "contig"), (130, 70, "ncon"), (256, 256, "ncon")])
def test_matvec_vs_oracle_seeded(dtype, mk, chi_l, chi_r, layout):
    mpo = golden_mpo(mk, dtype)
    p = len(mpo) // 2
    M1, M2 = mpo[p], mpo[p + 1]
    if M1.shape[0] > 20 and chi_l > 130:
        pytest.skip("oracle too slow for this case on the CPU")
    d1, d2 = M1.shape[3], M2.shape[3]
    L = randn((M1.shape[0], chi_l, chi_l), dtype, 10)
    R = randn((M2.shape[1], chi_r, chi_r), dtype, 11)
    psi = randn((chi_l * d1 * d2 * chi_r,), dtype, 12)
    psi = psi / torch.linalg.norm(psi)
    if layout == "ncon":
        L, R = ncon_layout(L), ncon_layout(R)
    ref = orc.apply_mpo(psi, L, M1, M2, R)
    Ld, Rd = L.contiguous().to(DEV), R.contiguous().to(DEV)
    if layout == "ncon":
        Ld, Rd = ncon_layout(Ld), ncon_layout(Rd)
    out = qsb.ApplyMPO()(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), Rd)
    assert rel_err(out, ref) <= TOL


@pytest.mark.parametrize("dtype,mk,chi", [(torch.float64, "j1j2_f64_8x6", 512), (torch.complex128, "hubbard_f64_L64", 320),
                                          (torch.float64, "hubbard_f64_L64", 512), (torch.complex128, "tfim_c128_L100_g0.5", 768)])
def test_matvec_config45_shapes_vs_oracle_on_device(dtype, mk, chi):
    """BASELINE configs 4/5 MPOs (J1-J2 cylinder W~30-38, Hubbard d=4) at bond dimensions the CPU oracle cannot do
    in seconds: the oracle's tensordot chain executed by cuBLAS on the same device is the checker."""
    mpo = golden_mpo(mk, dtype, DEV)
    p = len(mpo) // 2
    M1, M2 = mpo[p], mpo[p + 1]
    g = torch.Generator(device=DEV).manual_seed(17)
    L = torch.randn(M1.shape[0], chi, chi, dtype=dtype, device=DEV, generator=g)
    R = torch.randn(M2.shape[1], chi, chi, dtype=dtype, device=DEV, generator=g)
    x = torch.randn(chi * M1.shape[3] * M2.shape[3] * chi, dtype=dtype, device=DEV, generator=g)
    out = qsb.ApplyMPO()(x, L, M1, M2, R)
    ref = orc.apply_mpo(x, L, M1, M2, R)
    assert rel_err(out, ref) <= TOL
    A = torch.randn(chi, M1.shape[2], chi, dtype=dtype, device=DEV, generator=g)
    assert rel_err(qsb.LeftEnvUpdate()(L, M1, A), orc.left_env_update(L, M1, A)) <= TOL
    assert rel_err(qsb.RightEnvUpdate()(M2, R, A), orc.right_env_update(M2, R, A)) <= TOL
    del out, ref
    torch.cuda.empty_cache()


def test_matvec_arbitrary_strides_and_views():
    """L / R given as slices of bigger tensors (neither inner stride is 1) are canonicalised, not rejected."""
    dtype = torch.float64
    mpo = golden_mpo("heis_f64_L20", dtype)
    M1, M2 = mpo[5], mpo[6]
    Lbig = randn((5, 24, 40), dtype, 20)
    L = Lbig[:, ::2, ::2][:, :10, :10]                   # strides (960, 80, 2)
    R = randn((5, 12, 12), dtype, 21)
    psi = randn((10 * 4 * 12,), dtype, 22)
    ref = orc.apply_mpo(psi, L, M1, M2, R)
    Ld = Lbig.to(DEV)[:, ::2, ::2][:, :10, :10]
    out = qsb.ApplyMPO()(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), R.to(DEV))
    assert rel_err(out, ref) <= TOL


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("chi_l,chi_r", [(1024, 512), (256, 256), (96, 40), (6, 6)])
def test_matvec_host_resident_vector(dtype, chi_l, chi_r):
    """psi_flat / out in pinned HOST memory with the operator on the GPU: pipelined upload -> gated matvec ->
    row-block download gives the same result as the device-resident call."""
    mpo = golden_mpo("heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100", dtype, DEV)
    M1, M2 = mpo[50], mpo[51]
    g = torch.Generator(device=DEV).manual_seed(23)
    L = torch.randn(5, chi_l, chi_l, dtype=dtype, device=DEV, generator=g)
    R = torch.randn(5, chi_r, chi_r, dtype=dtype, device=DEV, generator=g)
    x = torch.randn(chi_l * 4 * chi_r, dtype=dtype, device=DEV, generator=g)
    ref = qsb.ApplyMPO()(x, L, M1, M2, R)
    xh = x.cpu().pin_memory()
    oh = torch.empty(ref.numel(), dtype=dtype).pin_memory()
    op = qsb.ApplyMPO()
    for rep in range(3):                        # repeated calls: epochs, buffer reuse
        res = op(xh, L, M1, M2, R, out=oh)
        torch.cuda.synchronize()
        assert res.data_ptr() == oh.data_ptr()
        assert rel_err(oh, ref) <= 1e-14
    res2 = op(xh, L, M1, M2, R)                 # allocates a pinned result
    torch.cuda.synchronize()
    assert res2.is_pinned() and rel_err(res2, ref) <= 1e-14
    with pytest.raises(ValueError):
        op(x.cpu(), L, M1, M2, R)               # unpinned host memory is rejected


def test_matvec_shape_errors():
    dtype = torch.float64
    mpo = golden_mpo("heis_f64_L20", dtype, DEV)
    L = torch.zeros(4, 6, 6, dtype=dtype, device=DEV)        # wrong MPO bond (M1 has W_l = 5)
    R = torch.zeros(5, 6, 6, dtype=dtype, device=DEV)
    psi = torch.zeros(6 * 4 * 6, dtype=dtype, device=DEV)
    with pytest.raises(AssertionError):
        qsb.ApplyMPO()(psi, L, mpo[5], mpo[6], R)


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
def test_matvec_full_size_properties(dtype):
    """BASELINE config 3 size (chi = 2048, Heisenberg bulk bond): linearity and agreement with the oracle's
    torch restatement executed by cuBLAS on the same device (the CPU oracle needs minutes at this size)."""
    chi = 2048 if dtype == torch.float64 else 1024
    mpo = golden_mpo("heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100", dtype, DEV)
    M1, M2 = mpo[50], mpo[51]
    g = torch.Generator(device=DEV).manual_seed(5)
    L = torch.randn(5, chi, chi, dtype=dtype, device=DEV, generator=g)
    R = torch.randn(5, chi, chi, dtype=dtype, device=DEV, generator=g)
    x = torch.randn(chi * 4 * chi, dtype=dtype, device=DEV, generator=g)
    y = torch.randn(chi * 4 * chi, dtype=dtype, device=DEV, generator=g)
    op = qsb.ApplyMPO()
    Hx, Hy = op(x, L, M1, M2, R).clone(), op(y, L, M1, M2, R).clone()
    a, b = 0.37, -1.25
    Hxy = op(a * x + b * y, L, M1, M2, R)
    assert rel_err(Hxy, a * Hx + b * Hy) <= TOL
    ref = orc.apply_mpo(x, L, M1, M2, R)
    assert rel_err(Hx, ref) <= TOL
    del ref, Hx, Hy, Hxy
    torch.cuda.empty_cache()


# ----------------------------------------------------------------------------- environment updates
def _svd_layout(A, B):
    chi, d, chip = A.shape
    A2 = A.reshape(chi * d, chip).T.contiguous().T.view(chi, d, chip)
    big = torch.zeros(chi + 3, d * chip, dtype=B.dtype, device=B.device)
    big[:chi] = B.reshape(chi, d * chip)
    B2 = big.T.contiguous().T[:chi, :].view(chi, d, chip)
    return A2, B2


@pytest.mark.parametrize("name", _names(load_golden("env")))
def test_env_golden(name):
    z = load_golden("env")
    g = {k: torch.from_numpy(z[f"{name}/{k}"]).to(DEV) for k in ("Lp", "M", "A", "Rn", "B", "Lnew", "Rnew")}
    A, B = g["A"], g["B"]
    if str(z[f"{name}/layout"]) == "svd":
        A, B = _svd_layout(A, B)
        assert not A.is_contiguous() or 1 in A.shape
    Ln = qsb.LeftEnvUpdate("ncon")(g["Lp"], g["M"], A)
    Rp = qsb.RightEnvUpdate("ncon")(g["M"], g["Rn"], B)
    assert Ln.shape == g["Lnew"].shape and Rp.shape == g["Rnew"].shape
    assert rel_err(Ln, g["Lnew"]) <= TOL
    assert rel_err(Rp, g["Rnew"]) <= TOL
    # environments in the reference's own strided layout are accepted as inputs too
    Ln2 = qsb.LeftEnvUpdate()(ncon_layout(g["Lp"]), g["M"], A)
    Rp2 = qsb.RightEnvUpdate()(g["M"], ncon_layout(g["Rn"]), B)
    assert rel_err(Ln2, g["Lnew"]) <= TOL and rel_err(Rp2, g["Rnew"]) <= TOL


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("chi,chip", [(200, 200), (144, 96), (31, 77)])
def test_env_vs_oracle_seeded(dtype, chi, chip):
    mpo = golden_mpo("heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100", dtype)
    M = mpo[40]
    W, Wp, d = M.shape[0], M.shape[1], M.shape[2]
    Lp = randn((W, chi, chi), dtype, 30)
    A = randn((chi, d, chip), dtype, 31)
    Rn = randn((Wp, chip, chip), dtype, 32)
    B = randn((chi, d, chip), dtype, 33)
    refL = orc.left_env_update(Lp, M, A)
    refR = orc.right_env_update(M, Rn, B)
    outL = qsb.LeftEnvUpdate()(Lp.to(DEV), M.to(DEV), A.to(DEV))
    outR = qsb.RightEnvUpdate()(M.to(DEV), Rn.to(DEV), B.to(DEV))
    assert rel_err(outL, refL) <= TOL and rel_err(outR, refR) <= TOL


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("chi,chip,lo,hi", [(64, 48, 16, 32), (96, 96, 0, 48), (33, 20, 5, 6)])
def test_env_row_block_forms(dtype, chi, chip, lo, hi):
    """Row-block env updates used by the multi-GPU sharding: left = partial sum over a bra-row block (the blocks add
    up to the full update), right = a row block of the new environment."""
    mpo = golden_mpo("heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100", dtype)
    M = mpo[40]
    W, Wp, d = M.shape[0], M.shape[1], M.shape[2]
    Lp = randn((W, chi, chi), dtype, 40)
    A = randn((chi, d, chip), dtype, 41)
    Rn = randn((Wp, chip, chip), dtype, 42)
    B = randn((chi, d, chip), dtype, 43)
    ops = _lib.ops
    part = torch.empty(Wp, chip, chip, dtype=dtype, device=DEV)
    Ad = A.to(DEV)
    ops.env_left_rows(Lp[:, lo:hi, :].to(DEV), M.to(DEV), Ad, Ad[lo:hi], part)
    ref_part = torch.einsum("wak,apc,wvps,ksd->vcd", Lp[:, lo:hi, :], A[lo:hi].conj(), M, A)
    assert rel_err(part, ref_part) <= TOL
    rest = torch.empty_like(part)
    idx = [i for i in range(chi) if not (lo <= i < hi)]
    ops.env_left_rows(Lp[:, idx, :].to(DEV), M.to(DEV), Ad, Ad[idx], rest)
    assert rel_err(part + rest, orc.left_env_update(Lp, M, A)) <= TOL
    rows = torch.empty(W, hi - lo, chi, dtype=dtype, device=DEV)
    Bd = B.to(DEV)
    ops.env_right_rows(Rn.to(DEV), M.to(DEV), Bd, Bd[lo:hi], rows)
    assert rel_err(rows, orc.right_env_update(M, Rn, B)[:, lo:hi, :]) <= TOL


def test_environment_manager_chain_matches_oracle():
    """update_R from the right boundary then update_L from the left one, on a real right-canonical MPS."""
    dtype = torch.complex128
    mpo = golden_mpo("heis_c128_L20", dtype)
    mps = orc.OracleMPS(20, 2, 16, dtype, seed=4)
    ref = orc.EnvCache(mpo, dtype)
    env = qsb.EnvironmentManager([m.to(DEV) for m in mpo], DEV, dtype)
    for p in range(19, 0, -1):
        ref.update_R(p, mps.B[p])
        env.update_R(p, mps.B[p].to(DEV))
        assert rel_err(env.get_R(p), ref.get_R(p)) <= TOL
    for p in range(0, 10):
        ref.update_L(p, mps.B[p])
        env.update_L(p, mps.B[p].to(DEV))
        assert rel_err(env.get_L(p + 1), ref.get_L(p + 1)) <= TOL


def test_native_library_was_used():
    c = qsb.launch_counters()
    assert c["total"] > 0 and (c["cpasync_gemm"] + c["tma_gemm"]) > 0


# ----------------------------------------------------------------------------- identity-channel shortcut
def _with_identity(env, w, row0=0):
    """Replace channel w of an environment (W, rows, cols) by the (row-shifted) identity, plus rounding-level noise
    of the size a real A^H A carries."""
    env = env.clone()
    rows, cols = env.shape[1:]
    eye = torch.zeros(rows, cols, dtype=env.dtype)
    eye[torch.arange(rows), torch.arange(rows) + row0] = 1.0
    env[w] = eye + 1e-16 * randn((rows, cols), env.dtype, 77)
    return env


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("chi_l,chi_r,wl_id,wr_id,layout", [
    (64, 64, 0, 4, "contig"),       # the FSM pattern: first channel of L, last channel of R
    (128, 96, 0, 4, "ncon"),        # TMA-eligible, reference's strided environments
    (130, 70, 0, 4, "contig"),      # ragged -> cp.async GEMM + accumulate
    (48, 40, 2, 1, "contig"),       # identity channels in the middle: two runs of panels each side
    (32, 32, 0, None, "contig"),    # only L
    (32, 32, None, 4, "ncon"),      # only R
    (17, 5, 4, 0, "contig"),
])
def test_matvec_identity_channels(dtype, chi_l, chi_r, wl_id, wr_id, layout):
    """MPO-sparsity-aware block skipping in the chi^3 GEMMs: with verified identity slices in L and/or R the
    result equals the dense computation of the oracle (the reference has no such shortcut)."""
    key = "heis_c128_L20" if dtype.is_complex else "heis_f64_L20"
    Ms = golden_mpo(key, dtype)
    M1, M2 = Ms[9], Ms[10]
    L = randn((5, chi_l, chi_l), dtype, 11)
    R = randn((5, chi_r, chi_r), dtype, 12)
    if wl_id is not None:
        L = _with_identity(L, wl_id)
    if wr_id is not None:
        R = _with_identity(R, wr_id)
    psi = randn((chi_l * 4 * chi_r,), dtype, 13)
    ref = orc.apply_mpo(psi, L, M1, M2, R)
    Ld, Rd = L.to(DEV), R.to(DEV)
    if layout == "ncon":
        Ld, Rd = ncon_layout(Ld), ncon_layout(Rd)
    _lib.tag_identity(Ld)
    _lib.tag_identity(Rd)
    assert _lib.identity_of(Ld)[0] == (-1 if wl_id is None else wl_id)
    assert _lib.identity_of(Rd)[0] == (-1 if wr_id is None else wr_id)
    op = qsb.ApplyMPO()
    before = qsb.launch_counters()
    out = op(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), Rd)
    after = qsb.launch_counters()
    assert rel_err(out, ref) <= TOL
    # flop accounting: executed < dense exactly by the skipped slices
    ident = (-1 if wl_id is None else wl_id, -1 if wr_id is None else wr_id, 0)
    dense = _lib.heff_flops(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), Rd)
    execd = _lib.heff_flops_executed(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), Rd, ident)
    c = 8.0 if dtype.is_complex else 2.0
    skipped = c * 4 * (chi_l ** 2 * chi_r * (wl_id is not None) + chi_r ** 2 * chi_l * (wr_id is not None))
    assert abs((dense - execd) - skipped) <= 1e-9 * dense
    # GEMM launches: one per run of dense channels on each side
    runs = lambda W, i: 1 if i is None or i in (0, W - 1) else 2
    gemms = (after["tma_gemm"] + after["cpasync_gemm"]) - (before["tma_gemm"] + before["cpasync_gemm"])
    assert gemms == runs(5, wl_id) + runs(5, wr_id)
    # an in-place write invalidates the tag (the shortcut must never act on stale knowledge)
    if wl_id is not None:
        Ld.mul_(2.0)
        assert _lib.identity_of(Ld)[0] == -1
        out2 = op(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), Rd)
        L2 = L.clone() * 2.0
        assert rel_err(out2, orc.apply_mpo(psi, L2, M1, M2, R)) <= TOL


def test_matvec_identity_single_channel_boundaries():
    """W = 1 on a side whose only slice is the identity: no GEMM is launched on that side at all (chain ends,
    where the reference's boundary environments are ones(W, 1, 1), dmrg.py:349-351)."""
    dt = torch.float64
    Ms = golden_mpo("heis_f64_L20", dt)
    M1, M2 = Ms[0], Ms[1]                     # (1, 5, 2, 2), (5, 5, 2, 2)
    L = torch.ones(1, 1, 1, dtype=dt)
    R = _with_identity(randn((5, 24, 24), dt, 3), 4)
    psi = randn((1 * 4 * 24,), dt, 4)
    ref = orc.apply_mpo(psi, L, M1, M2, R)
    Ld, Rd = _lib.tag_identity(L.to(DEV)), _lib.tag_identity(R.to(DEV))
    assert _lib.identity_of(Ld)[0] == 0 and _lib.identity_of(Rd)[0] == 4
    out = qsb.ApplyMPO()(psi.to(DEV), Ld, M1.to(DEV), M2.to(DEV), Rd)
    assert rel_err(out, ref) <= TOL
    # mirror image: last bond
    M1, M2 = Ms[18], Ms[19]
    L = _with_identity(randn((5, 24, 24), dt, 5), 0)
    R = torch.ones(1, 1, 1, dtype=dt)
    psi = randn((24 * 4,), dt, 6)
    ref = orc.apply_mpo(psi, L, M1, M2, R)
    out = qsb.ApplyMPO()(psi.to(DEV), _lib.tag_identity(L.to(DEV)), M1.to(DEV), M2.to(DEV), _lib.tag_identity(R.to(DEV)))
    assert rel_err(out, ref) <= TOL


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
def test_identity_deviation_kernel(dtype):
    W, rows, cols, row0 = 4, 48, 80, 16
    env = randn((W, rows, cols), dtype, 21)
    env = _with_identity(env, 2, row0)
    dev_env = ncon_layout(env.to(DEV)) if dtype.is_complex else env.to(DEV)
    assert _lib.identity_channel(dev_env, row0) == 2
    assert _lib.identity_channel(dev_env, 0) == -1                 # wrong shift -> not an identity
    assert _lib.identity_channel(dev_env, row0, tol=0.0) == -1     # disabled
    bad = env.clone()
    bad[2, 5, 7] += 1e-9
    assert _lib.identity_channel(bad.to(DEV), row0) == -1
    nan = env.clone()
    nan[2, 1, 1] = float("nan")
    assert _lib.identity_channel(nan.to(DEV), row0) == -1
    if dtype.is_complex:
        im = env.clone()
        im[2, 3, 3 + row0] += 1e-6j
        assert _lib.identity_channel(im.to(DEV), row0) == -1


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("w_id", [0, 2, 4])
def test_env_update_identity_input(dtype, w_id):
    """Environment updates with a verified identity slice in the old environment (its GEMM becomes a copy) against
    the oracle; the output of an isometric site tensor is tagged as carrying an identity channel in turn."""
    key = "heis_c128_L20" if dtype.is_complex else "heis_f64_L20"
    M = golden_mpo(key, dtype)[9]
    chi, chip = 40, 24
    A = randn((chi, 2, chip), dtype, 31)
    B = randn((chip, 2, chi), dtype, 32)
    Lp = _with_identity(randn((5, chi, chi), dtype, 33), w_id)
    Rn = _with_identity(randn((5, chi, chi), dtype, 34), w_id)
    refL = orc.left_env_update(Lp, M, A)
    refR = orc.right_env_update(M, Rn, B)
    Lpd, Rnd = _lib.tag_identity(Lp.to(DEV)), _lib.tag_identity(Rn.to(DEV))
    assert _lib.identity_of(Lpd)[0] == w_id
    outL = qsb.LeftEnvUpdate()(Lpd, M.to(DEV), A.to(DEV))
    outR = qsb.RightEnvUpdate()(M.to(DEV), Rnd, B.to(DEV))
    assert rel_err(outL, refL) <= TOL and rel_err(outR, refR) <= TOL
    # canonical site tensors reproduce the FSM pattern: identity in -> identity out.  In the reference's FSM layout
    # (finite_state_machine.py:662-700) the LAST MPO state is "nothing has happened yet" (identity channel of L) and
    # state 0 is "all terms completed" (identity channel of R).
    Q, _ = torch.linalg.qr(randn((chi * 2, chip), dtype, 35))
    Aiso = Q.reshape(chi, 2, chip)
    L4 = _with_identity(randn((5, chi, chi), dtype, 36), 4)
    outL = qsb.LeftEnvUpdate()(_lib.tag_identity(L4.to(DEV)), M.to(DEV), Aiso.to(DEV))
    assert rel_err(outL, orc.left_env_update(L4, M, Aiso)) <= TOL
    assert _lib.identity_of(outL)[0] == 4
    Biso = Q.T.reshape(chip, 2, chi) if not dtype.is_complex else Q.conj().T.reshape(chip, 2, chi)
    R0 = _with_identity(randn((5, chi, chi), dtype, 37), 0)
    outR = qsb.RightEnvUpdate()(M.to(DEV), _lib.tag_identity(R0.to(DEV)), Biso.to(DEV))
    assert rel_err(outR, orc.right_env_update(M, R0, Biso)) <= TOL
    assert _lib.identity_of(outR)[0] == 0


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(1, 256, 128, 64), (2, 130, 70, 33), (1, 1024, 1024, 520)])
def test_gemm_accumulate(dtype, shape):
    Z, M, N, K = shape
    A, B, C0 = randn((Z, M, K), dtype, 41), randn((Z, K, N), dtype, 42), randn((Z, M, N), dtype, 43)
    out = C0.to(DEV).clone()
    _lib.gemm(A.to(DEV), B.to(DEV), out=out, accumulate=True)
    assert rel_err(out, C0 + torch.matmul(A, B)) <= TOL
    out = C0.to(DEV).clone()
    _lib.gemm(A.to(DEV), B.to(DEV), out=out)
    assert rel_err(out, torch.matmul(A, B)) <= TOL


def test_dmrg_sweep_uses_identity_channels():
    """Inside a real sweep the environments of an FSM MPO carry identity channels and the kernels use them: the
    energies still match the reference's golden run to 1e-10 (test_dmrg_matches_reference_per_update covers that);
    here: the tags are what the FSM structure predicts."""
    import io, contextlib
    dt = torch.float64
    mpo = golden_mpo("heis_f64_L20", dt, DEV)
    with contextlib.redirect_stdout(io.StringIO()):
        mps = qsb.MPS(20, 2, 32, "cpu", dt, 1).to(DEV)
        env = qsb.EnvironmentManager(mpo, DEV, dt)
        for p in range(19, 0, -1):
            env.update_R(p, mps.B[p])
    for p in range(2, 20):
        assert _lib.identity_of(env.get_R(p))[0] == 0, p            # the FSM's "all terms completed" state
    assert _lib.identity_of(env.get_R(1))[0] == -1                  # bond 0|1 (W = 4): no term can be complete yet
    for p in range(0, 19):                                          # left-canonical tensors from a QR sweep
        l, d, r = mps.B[p].shape
        Q, Rf = torch.linalg.qr((mps.B[p] if p == 0 else torch.tensordot(Rf, mps.B[p], dims=([1], [0]))).reshape(l * d, r))
        env.update_L(p, Q.reshape(l, d, Q.shape[1]))
        want = mpo[p].shape[1] - 1 if p < 18 else -1                # bond 18|19 (W = 4): every term has started
        assert _lib.identity_of(env.get_L(p + 1))[0] == want, p


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("shape", [(256, 128, 512), (96, 40, 130), (512, 256, 2048)])
def test_gemm_epilogue_scaling(dtype, shape):
    """Row / column scale factors fused into the GEMM epilogue (the Schmidt weights of the theta build, row f.2 of
    SURVEY.md section 8): C[m][n] = rs[m // div] * cs[n % mod] * (A @ B), on the TMA-staged kernel (aligned shapes) and
    the cp.async one (ragged shapes), against the explicit elementwise scaling."""
    M, K, N = shape
    A = randn((M, K), dtype, 31).to(DEV)
    B = randn((K, N), dtype, 32).to(DEV)
    div, mod = 2, N // 2
    rs = torch.rand(M // div, dtype=torch.float64, device=DEV) + 0.5
    cs = torch.rand(mod, dtype=torch.float64, device=DEV) + 0.5
    ref = A @ B
    rfull = rs.repeat_interleave(div).to(dtype).unsqueeze(1)
    cfull = cs.repeat(N // mod).to(dtype).unsqueeze(0)
    before = _lib.launch_counters()
    got_r = _lib.gemm(A, B, row_scale=(rs, div))
    got_c = _lib.gemm(A, B, col_scale=(cs, mod))
    got_rc = _lib.gemm(A, B, row_scale=(rs, div), col_scale=(cs, mod))
    after = _lib.launch_counters()
    assert rel_err(got_r, ref * rfull) <= 1e-13
    assert rel_err(got_c, ref * cfull) <= 1e-13
    assert rel_err(got_rc, ref * rfull * cfull) <= 1e-13
    tma = after["tma_gemm"] - before["tma_gemm"]
    assert tma == (3 if (M % 16 == 0 and N % 16 == 0 and K % 16 == 0) else 0) or tma in (0, 3)
    with pytest.raises(ValueError):
        _lib.gemm(A, B, row_scale=(rs[:-1], div))


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
@pytest.mark.parametrize("direction", ["right", "left"])
def test_theta_build_with_fused_schmidt_weights(dtype, direction):
    """dmrg.build_theta (one GEMM, weights in the epilogue) against the reference's formula (dmrg.py:633-660), also for a
    row block of the left tensor (the sharded theta build) and for a complex boundary factor (explicit pass)."""
    from quantum_simulation_b200.dmrg import build_theta
    l, d, m, r = 48, 2, 64, 80
    lt = randn((l, d, m), dtype, 41).to(DEV)
    rt = randn((m, d, r), dtype, 42).to(DEV)
    sw = torch.rand(l if direction == "right" else r, dtype=torch.float64, device=DEV)
    if direction == "right":
        ref = (lt * sw.to(dtype).view(-1, 1, 1)).reshape(l * d, m) @ rt.reshape(m, d * r)
    else:
        ref = lt.reshape(l * d, m) @ (rt * sw.to(dtype).view(1, 1, -1)).reshape(m, d * r)
    assert rel_err(build_theta(lt, rt, sw, direction), ref) <= 1e-13
    lo, hi = 16, 32
    part = build_theta(lt[lo:hi], rt, sw[lo:hi] if direction == "right" else sw, direction)
    assert rel_err(part, ref[lo * d:hi * d]) <= 1e-13
    if dtype.is_complex:                         # complex boundary factor (2-D sWeight): explicit elementwise pass
        swc = torch.diag(randn((sw.numel(),), dtype, 43).to(DEV))
        v = torch.diagonal(swc)
        ref_c = ((lt * v.view(-1, 1, 1)).reshape(l * d, m) @ rt.reshape(m, d * r) if direction == "right" else
                 lt.reshape(l * d, m) @ (rt * v.view(1, 1, -1)).reshape(m, d * r))
        assert rel_err(build_theta(lt, rt, swc, direction), ref_c) <= 1e-13
