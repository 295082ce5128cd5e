"""CPU oracle for the DDDM hot path — TEST INFRASTRUCTURE ONLY.

Two pieces live here:

* ``dddm_oracle.c`` (bound below through ctypes): a plain-C, double-precision restatement of
  ``dddm/losses.py``, ``dddm/schedules.py`` and the update of ``dddm/sampling.py`` from the
  reference.  It is the checker the GPU parity tests compare the CUDA kernels with.
* ``torch_port.py``: a CPU PyTorch (eager + autograd) port of the same path, used as the timed
  CPU baseline in ``bench.py`` (``cpu_baseline`` / ``--impl reference``), because the reference's
  own CPU path is eager PyTorch and ``/root/reference`` does not exist on the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this
package.  ``ddm_b200`` never does: the product path fails loudly without its CUDA library.

Parity pinning: the reference ships no tests or golden vectors.  The oracle is pinned against
outputs of the reference's own functions captured in ``tests/golden/`` (generator script
``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile ``dddm_oracle.c`` with gcc (see ``oracle/Makefile``)."""
    src = os.path.join(_HERE, "dddm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "_build/liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_energy_terms.argtypes = [_dp, _dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_double,
                                          _dp, _dp, _dp, _dp]
        L.oracle_energy_terms_grad.argtypes = [_dp, _dp, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                               ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp, _dp]
        L.oracle_sigmoid_weight.argtypes = [_dp, ctypes.c_long, ctypes.c_double, _dp]
        L.oracle_energy_loss.argtypes = [_dp, _dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, _dp, _dp]
        L.oracle_forward_marginal.argtypes = [_dp, _dp, _dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, _dp, _dp]
        L.oracle_bridge_coeffs.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp]
        L.oracle_bridge_step.argtypes = [_dp, _dp, _dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_long,
                                         ctypes.c_long, _dp, _dp]
        for name in ("oracle_energy_terms", "oracle_energy_terms_grad", "oracle_sigmoid_weight",
                     "oracle_energy_loss", "oracle_forward_marginal", "oracle_bridge_coeffs",
                     "oracle_bridge_step"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _f64(a) -> np.ndarray:
    """Widen anything array-like (numpy, torch fp32/bf16/fp64 on any device) to contiguous float64."""
    if hasattr(a, "detach"):
        a = a.detach().to("cpu").double().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a: np.ndarray | None):
    return a.ctypes.data_as(_dp) if a is not None else None


def energy_terms(xhat, x0, beta: float):
    """(conf, inter, row_conf[B], row_inter[B]) of ``generalized_energy_terms`` (dddm/losses.py:5-25)."""
    xh, c = _f64(xhat), _f64(x0)
    B, m, D = xh.shape
    assert c.shape == (B, D)
    conf, inter = ctypes.c_double(), ctypes.c_double()
    rc, ri = np.empty(B), np.empty(B)
    lib().oracle_energy_terms(_p(xh), _p(c), B, m, D, float(beta), ctypes.byref(conf), ctypes.byref(inter),
                              _p(rc), _p(ri))
    return conf.value, inter.value, rc, ri


def energy_terms_grad(xhat, x0, beta: float, g_conf: float, g_inter: float, want_x0: bool = False):
    """Closed-form d(g_conf*conf + g_inter*inter)/d(xhat[, x0])."""
    xh, c = _f64(xhat), _f64(x0)
    B, m, D = xh.shape
    g = np.empty_like(xh)
    g0 = np.empty_like(c) if want_x0 else None
    lib().oracle_energy_terms_grad(_p(xh), _p(c), B, m, D, float(beta), float(g_conf), float(g_inter), _p(g), _p(g0))
    return (g, g0) if want_x0 else g


def sigmoid_weight(t, bias: float = 0.0) -> np.ndarray:
    """w(t) of dddm/losses.py:28-35."""
    tt = _f64(t).reshape(-1)
    w = np.empty_like(tt)
    lib().oracle_sigmoid_weight(_p(tt), tt.size, float(bias), _p(w))
    return w


def energy_loss(xhat, x0, beta: float, lam: float, weight: float, want_grad: bool = True):
    """(loss, conf, inter, dloss/dxhat) of dddm/training.py:84-85 for a given mean weight."""
    xh, c = _f64(xhat), _f64(x0)
    B, m, D = xh.shape
    out = np.empty(3)
    g = np.empty_like(xh) if want_grad else None
    lib().oracle_energy_loss(_p(xh), _p(c), B, m, D, float(beta), float(lam), float(weight), _p(out), _p(g))
    return out[0], out[1], out[2], g


def forward_marginal(x0, t, eps, m: int = 0):
    """(xt [B,D], xt_rep [B*m,D] or None) — dddm/schedules.py:17-25 + the expansion of training.py:70."""
    c, e = _f64(x0), _f64(eps)
    B = c.shape[0]
    c2, e2 = c.reshape(B, -1), e.reshape(B, -1)
    D = c2.shape[1]
    tt = _f64(t).reshape(-1)
    assert tt.size == B
    xt = np.empty_like(c2)
    rep = np.empty((B * m, D)) if m > 0 else None
    lib().oracle_forward_marginal(_p(c2), _p(tt), _p(e2), B, D, m, _p(xt), _p(rep))
    return xt, rep


def bridge_coeffs(s: float, t: float, eps_churn: float = 1.0):
    """(c_xt, c_x0, std) of dddm/schedules.py:45-77."""
    coef = np.empty(3)
    lib().oracle_bridge_coeffs(float(s), float(t), float(eps_churn), _p(coef))
    return tuple(coef)


def bridge_step(x, x0hat, z, s, t, eps_churn: float = 1.0):
    """(x_next, mu) of dddm/sampling.py:29-31; ``s``/``t`` scalars or per-sample vectors."""
    xx, xh = _f64(x), _f64(x0hat)
    N = xx.shape[0]
    x2, h2 = xx.reshape(N, -1), xh.reshape(N, -1)
    D = x2.shape[1]
    zz = _f64(z).reshape(N, D) if z is not None else None
    ss, tt = _f64(s).reshape(-1), _f64(t).reshape(-1)
    vec = int(ss.size > 1 or tt.size > 1)
    if vec:
        ss = np.ascontiguousarray(np.broadcast_to(ss, (N,)))
        tt = np.ascontiguousarray(np.broadcast_to(tt, (N,)))
    out, mu = np.empty_like(x2), np.empty_like(x2)
    lib().oracle_bridge_step(_p(x2), _p(h2), _p(zz), _p(ss), _p(tt), vec, float(eps_churn), N, D, _p(out), _p(mu))
    return out.reshape(xx.shape), mu.reshape(xx.shape)


def rbf_mmd2(x, y, sigma: float = 1.0):
    """numpy fp64 restatement of dddm/metrics.py:140-163 (unbiased MMD^2, RBF kernel, sigma fixed): the Gram-form
    squared distances a2 + b2 - 2 a.b (:143-146), exp(-d2 / (2 sigma^2)), off-diagonal means for xx / yy (:157-160),
    full mean for xy (:161).  Returns (mmd2, kxx, kyy, kxy)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    n, m = x.shape[0], y.shape[0]
    if n < 2 or m < 2:
        raise ValueError("Need at least two samples per set to compute MMD")
    gamma = 1.0 / (2.0 * sigma**2)

    def pdist2(a, b):
        return (a * a).sum(-1)[:, None] + (b * b).sum(-1)[None, :] - 2.0 * (a @ b.T)

    kxx_full = np.exp(-gamma * pdist2(x, x))
    kyy_full = np.exp(-gamma * pdist2(y, y))
    kxx = (kxx_full.sum() - np.trace(kxx_full)) / (n * (n - 1))
    kyy = (kyy_full.sum() - np.trace(kyy_full)) / (m * (m - 1))
    kxy = np.exp(-gamma * pdist2(x, y)).mean()
    return kxx + kyy - 2.0 * kxy, kxx, kyy, kxy
