"""qbm_b200 -- B200-native (sm_100a) implementation of the sampling-and-training hot path of
Mark-Seebode/QBM-Image-Classification.  Import name: ``qbm_b200`` (see ``qbm_b200.py`` at the repo
root; the directory name carries a hyphen and is not itself importable).

Public surface (mirrors the reference's call boundary):

* ``B200SASampler``            drop-in for ``src/qubo/sampler.py::LocalSASampler``
* ``shims.install()``          ``dimod`` / ``neal`` duck types for ``Disc_QBM`` (faster_dqbm.py)
* ``sa_sample / qubo_energies / phase_stats``  batched device-level calls over the C ABI
"""
from . import _lib, conv_deep_qbm, disc_qbm, dist, ising, rbm, sampler, shims  # noqa: F401
from .conv_deep_qbm import ConvDeepQBM  # noqa: F401
from .rbm import B200ClassificationRBM, gemm_tf32  # noqa: F401
from .disc_qbm import DiscQBM  # noqa: F401
from .sampler import (B200SASampler, SAResult, phase_stats, qubo_energies, qubo_to_ising_device,  # noqa: F401
                      sa_sample, sample_qubo_batch)

__all__ = ["B200SASampler", "DiscQBM", "ConvDeepQBM", "B200ClassificationRBM", "gemm_tf32", "SAResult", "sa_sample", "qubo_energies", "phase_stats", "qubo_to_ising_device",
           "sample_qubo_batch", "ising", "shims", "sampler"]
