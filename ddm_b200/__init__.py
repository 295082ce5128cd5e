"""ddm_b200 — B200-native (sm_100a) hot path of Distributional Diffusion Models.

Hand-written CUDA kernels behind a C ABI (``include/dddm_b200.h``), exposed with the same
function signatures as the reference's ``dddm.losses`` / ``dddm.schedules`` / ``dddm.sampling`` /
``dddm.training``.  CUDA-only: there is no CPU or eager fallback, and importing the compute
modules without the built library (``python -m ddm_b200.build``) fails loudly on first use.
"""
from .losses import generalized_energy_terms, sigmoid_weight
from .metrics import rbf_mmd2
from .patch import patch_reference, unpatch_reference
from .sampling import sample_dddm, sample_dddm_chunked, sample_dddm_sharded
from .schedules import alpha_sigma, forward_marginal_sample, gaussian_bridge_mu_sigma
from .training import DeferredMetrics, TrainConfig, distributional_training_step

__all__ = [
    "generalized_energy_terms", "sigmoid_weight", "alpha_sigma", "forward_marginal_sample",
    "gaussian_bridge_mu_sigma", "sample_dddm", "sample_dddm_sharded", "sample_dddm_chunked", "distributional_training_step",
    "TrainConfig", "DeferredMetrics", "patch_reference", "unpatch_reference", "rbf_mmd2",
]
