"""Test-only CPU compute engine for the sharded DMRG driver (gloo tests of the multi-rank HOST logic): the same
methods as quantum_simulation_b200.parallel.CudaEngine, built on the oracle / plain torch."""
import torch

from oracle import dmrg_oracle as orc
from vecops_torch import TorchVectorOps


class CpuEngine:
    def __init__(self):
        self.vector_ops = TorchVectorOps()

    @staticmethod
    def apply_mpo(v, L, M1, M2, R, out=None):
        return orc.apply_mpo(v, L, M1, M2, R, out=out)

    @staticmethod
    def matmul(a, b):
        return torch.matmul(a, b)

    @staticmethod
    def env_left(Lp, M, A):
        return orc.left_env_update(Lp, M, A).contiguous()

    @staticmethod
    def env_right(M, Rn, B):
        return orc.right_env_update(M, Rn, B).contiguous()

    @staticmethod
    def env_left_rows(Lp_rows, M, A, A_rows):
        return torch.einsum("wak,apc,wvps,ksd->vcd", Lp_rows, A_rows.conj(), M, A).contiguous()

    @staticmethod
    def env_right_rows(M, Rn, B, B_rows):
        return torch.einsum("wvps,vxb,ksb,apx->wak", M, Rn, B, B_rows.conj()).contiguous()

    @staticmethod
    def split(theta2d, maxdim, cutoff, **kw):
        from quantum_simulation_b200.svd import two_site_split
        return two_site_split(theta2d, maxdim, cutoff, **kw)

    @staticmethod
    def lanczos(v, op, args, **kw):
        return orc.lanczos_ground_state(v, op, args, **kw)
