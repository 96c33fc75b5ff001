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
