"""Test-only vector backend with the semantics of the qsb_lz_* kernels, in plain torch (runs on the CPU).
Used by the gloo multi-rank tests to exercise the HOST logic of the sharded Lanczos without a GPU; the product
path (quantum_simulation_b200.eigensolver.CudaVectorOps) never imports this."""
import torch


class TorchVectorOps:
    device_type = "cpu"

    def make_scratch(self, device):
        return torch.zeros(1, device=device)

    @staticmethod
    def _rows(X, m, ldx, n):
        return torch.as_strided(X, (m, n), (max(ldx, 0) if m > 1 else n, 1))

    def dot(self, X, m, ldx, w, n, scal, out_idx, flags, scratch):
        Xr = self._rows(X, m, ldx, n)
        v = (Xr.conj().to(torch.complex128) @ w.to(torch.complex128))
        re, im = v.real.clone(), v.imag.clone()
        if flags & 1:
            im.zero_()
        if flags & 2:
            re = re.sqrt()
            im.zero_()
        scal[out_idx:out_idx + m, 0] = re
        scal[out_idx:out_idx + m, 1] = im

    def axpy_norm(self, X, m, ldx, w, n, scal, coef_idx, norm_idx, flags, scratch):
        Xr = self._rows(X, m, ldx, n)
        c = torch.complex(scal[coef_idx:coef_idx + m, 0], scal[coef_idx:coef_idx + m, 1])
        upd = (c[:, None] * Xr.to(torch.complex128)).sum(0)
        w.sub_(upd.to(w.dtype) if w.is_complex() else upd.real.to(w.dtype))
        if norm_idx >= 0:
            s = float((w.abs().to(torch.float64) ** 2).sum())
            scal[norm_idx, 0] = s if (flags & 4) else s ** 0.5
            scal[norm_idx, 1] = 0.0

    def scale(self, src, dst, n, scal, idx):
        dst.copy_(src / scal[idx, 0].to(src.dtype))

    def combine(self, X, m, ldx, y, dst, n, scal, norm_idx, flags, scratch):
        Xr = self._rows(X, m, ldx, n)
        yy = torch.tensor(y, dtype=torch.float64)
        dst.copy_((yy.to(Xr.dtype)[:, None] * Xr).sum(0))
        if norm_idx >= 0:
            s = float((dst.abs().to(torch.float64) ** 2).sum())
            scal[norm_idx, 0] = s if (flags & 4) else s ** 0.5
            scal[norm_idx, 1] = 0.0
