"""quantum_simulation_b200: B200-native (sm_100a) two-site DMRG inner loop.

Drop-in for the hot path of ToelUl/quantum-simulation: the H_eff matvec, the left/right environment updates and
the Lanczos vector operations run as hand-written CUDA kernels (csrc/) behind the reference's own Python API
(``ApplyMPO``, ``LeftEnvUpdate``, ``RightEnvUpdate``, ``EnvironmentManager``, ``lanczos_ground_state``, ``DMRG``,
``Sweeps``, ``MPS``).  The kernels are reached through ``torch.ops.qsb200.*`` custom ops over the C ABI declared in
``include/qsb200.h``.  There is no CPU fallback: importing works anywhere, calling needs a CUDA device and the
built ``lib/libqsb200.so``.
"""
from . import _lib
from ._lib import launch_counters, load as load_library
from .mps import MPS, random_mps
from .eigensolver import lanczos_ground_state, clear_workspaces
from .observables import (mpo_expectation, energy_expectation, energy_variance, local_expectation, two_point_correlation,
                          string_order_parameter, many_body_polarization, structure_factor, correlation_length,
                          ground_state_fidelity, entanglement_entropy, entanglement_spectrum, product_expectation,
                          ground_state_energy, energy_density, excitation_gap, truncation_errors)
from .dmrg import (ApplyMPO, LeftEnvUpdate, RightEnvUpdate, EnvironmentManager, Sweeps, DMRG,
                   set_contract_backend)

__all__ = ["MPS", "random_mps", "lanczos_ground_state", "clear_workspaces", "ApplyMPO", "LeftEnvUpdate",
           "RightEnvUpdate", "EnvironmentManager", "Sweeps", "DMRG", "set_contract_backend", "launch_counters",
           "load_library", "mpo_expectation", "energy_expectation", "energy_variance", "local_expectation",
           "two_point_correlation", "string_order_parameter", "many_body_polarization", "structure_factor",
           "correlation_length", "ground_state_fidelity", "entanglement_entropy", "entanglement_spectrum",
           "product_expectation", "ground_state_energy", "energy_density", "excitation_gap", "truncation_errors"]
