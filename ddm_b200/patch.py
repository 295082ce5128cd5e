"""Run the reference's unmodified scripts on the CUDA kernels.

The reference binds its hot-path functions by name at import time
(``dddm/training.py:10,12``: ``from .losses import generalized_energy_terms, sigmoid_weight`` /
``from .schedules import forward_marginal_sample``; ``dddm/sampling.py:5``:
``from .schedules import gaussian_bridge_mu_sigma``), so replacing ``dddm.losses.*`` alone has no
effect; the names have to be rebound inside the modules that imported them.
"""
from __future__ import annotations

import importlib
from types import ModuleType
from typing import Optional


def patch_reference(dddm: Optional[ModuleType] = None, *, whole_step: bool = True, whole_sampler: bool = True,
                    metrics: bool = True) -> dict:
    """Rebind the reference package ``dddm`` (already importable) onto ``ddm_b200``.

    * always: ``dddm.losses.{generalized_energy_terms, sigmoid_weight}``,
      ``dddm.schedules.{forward_marginal_sample, gaussian_bridge_mu_sigma}`` and the copies of those
      names inside ``dddm.training`` / ``dddm.sampling``;
    * ``whole_step``: also ``distributional_training_step`` (fused K1 path) in ``dddm.training`` and ``dddm``;
    * ``whole_sampler``: also ``sample_dddm`` in ``dddm.sampling`` and ``dddm``;
    * ``metrics``: also ``rbf_mmd2`` in ``dddm.metrics`` and ``dddm`` (``dddm/metrics.py:140``, used at ``run_example.py:101``).

    Returns {qualified name: original object} so the caller can undo it with :func:`unpatch_reference`.
    Scripts that did ``from dddm import sample_dddm`` before patching keep the old binding — patch first.
    """
    from . import losses, sampling, schedules, training

    if dddm is None:
        dddm = importlib.import_module("dddm")
    mods = {name: importlib.import_module(f"{dddm.__name__}.{name}") for name in ("losses", "schedules", "training",
                                                                                 "sampling")}
    plan = [
        (mods["losses"], "generalized_energy_terms", losses.generalized_energy_terms),
        (mods["losses"], "sigmoid_weight", losses.sigmoid_weight),
        (mods["schedules"], "forward_marginal_sample", schedules.forward_marginal_sample),
        (mods["schedules"], "gaussian_bridge_mu_sigma", schedules.gaussian_bridge_mu_sigma),
        (mods["training"], "generalized_energy_terms", losses.generalized_energy_terms),
        (mods["training"], "sigmoid_weight", losses.sigmoid_weight),
        (mods["training"], "forward_marginal_sample", schedules.forward_marginal_sample),
        (mods["sampling"], "gaussian_bridge_mu_sigma", schedules.gaussian_bridge_mu_sigma),
    ]
    if whole_step:
        plan += [(mods["training"], "distributional_training_step", training.distributional_training_step),
                 (dddm, "distributional_training_step", training.distributional_training_step)]
    if whole_sampler:
        plan += [(mods["sampling"], "sample_dddm", sampling.sample_dddm), (dddm, "sample_dddm", sampling.sample_dddm)]
    if metrics:
        from . import metrics as _metrics

        plan += [(importlib.import_module(f"{dddm.__name__}.metrics"), "rbf_mmd2", _metrics.rbf_mmd2),
                 (dddm, "rbf_mmd2", _metrics.rbf_mmd2)]
    saved = {}
    for mod, name, new in plan:
        saved[(mod, name)] = getattr(mod, name, None)
        setattr(mod, name, new)
    return saved


def unpatch_reference(saved: dict) -> None:
    for (mod, name), old in saved.items():
        if old is None:
            delattr(mod, name)
        else:
            setattr(mod, name, old)
