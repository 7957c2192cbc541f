"""hommx_b200 -- the hot path of flxrcz/hommx (FE-HMM micro cell problems + macro stiffness
assembly) on NVIDIA B200 (sm_100a).  Host code is Python, the kernels are hand-written CUDA
reached through the C ABI of ``libhmx.so`` (include/hmx.h).  There is no CPU fallback."""
from . import fem, mesh, ufl  # noqa: F401
from .hmm import (  # noqa: F401
    BaseHMM,
    LinearElasticityHMM,
    LinearElasticityStratifiedHMM,
    PoissonHMM,
    PoissonPeriodicHMM,
    PoissonStratifiedHMM,
)

__all__ = [
    "PoissonHMM",
    "PoissonStratifiedHMM",
    "LinearElasticityHMM",
    "LinearElasticityStratifiedHMM",
    "PoissonPeriodicHMM",
    "BaseHMM",
    "mesh",
    "fem",
    "ufl",
]
