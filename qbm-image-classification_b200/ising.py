"""Host-side (float64, numpy) part of the sampler boundary: what dimod and neal do in Python
before the native loop runs.  Kept on the host so that the beta rule, the sweeps-per-beta rule and
the initial-state generator stay one-line swaps:

* dense QUBO -> spin model: ``dimod.BQM(Q, "BINARY")`` + ``change_vartype(SPIN)``
  (call sites src/qubo/sampler.py:7-8,31; src/model/faster_dqbm.py:577,619)
* default beta range: neal 0.5.9 ``_default_ising_beta_range`` (legacy rule, neal/sampler.py:281)
* schedule: ``num_sweeps_per_beta = max(1, num_sweeps // 1000.0)``, ``np.geomspace(hot, cold, num_betas)``
* initial states: ``np.random.RandomState(seed).choice([-1, 1], size=(num_reads, n))``
  (dimod/core/initialized.py:207)

All functions accept a single problem ``[n, n]`` or a batch ``[B, n, n]``.
"""
from __future__ import annotations

import math

import numpy as np


def qubo_to_ising(Q: np.ndarray):
    """Returns ``(h [B,n] f64, J [B,n,n] f64 symmetric zero-diagonal, offset [B] f64)``."""
    Q = np.asarray(Q, dtype=np.float64)
    single = Q.ndim == 2
    if single:
        Q = Q[None]
    if Q.ndim != 3 or Q.shape[1] != Q.shape[2]:
        raise ValueError(f"QUBO must be [n, n] or [B, n, n], got {Q.shape}")
    n = Q.shape[1]
    a = np.diagonal(Q, axis1=1, axis2=2).copy()
    Bm = Q + np.transpose(Q, (0, 2, 1))
    idx = np.arange(n)
    Bm[:, idx, idx] = 0.0
    h = a / 2.0 + Bm.sum(axis=2) / 4.0
    offset = a.sum(axis=1) / 2.0 + Bm.sum(axis=(1, 2)) / 8.0
    Bm /= 4.0                      # J = b / 4 in place (the sums above are taken before the scaling, as before)
    return h, Bm, offset


def is_linear_only(Q: np.ndarray) -> np.ndarray:
    """``len(bqm.quadratic) == 0`` (src/qubo/sampler.py:10-11): no non-zero off-diagonal bias."""
    Q = np.asarray(Q, dtype=np.float64)
    if Q.ndim == 2:
        Q = Q[None]
    Bm = Q + np.transpose(Q, (0, 2, 1))
    idx = np.arange(Q.shape[1])
    Bm[:, idx, idx] = 0.0
    return ~np.any(Bm != 0.0, axis=(1, 2))


def default_beta_range(h: np.ndarray, J: np.ndarray) -> np.ndarray:
    """Legacy neal rule on SPIN biases; returns ``[B, 2]`` = (hot, cold)."""
    h = np.atleast_2d(np.asarray(h, dtype=np.float64))
    J = np.asarray(J, dtype=np.float64)
    if J.ndim == 2:
        J = J[None]
    absh = np.abs(h)
    absJ = np.abs(J)
    big = np.inf
    max_delta = (absh + absJ.sum(axis=2)).max(axis=1)
    min_h = np.where(absh != 0, absh, big).min(axis=1)
    absJ[absJ == 0.0] = big        # in place: absJ is a private temporary
    min_j = absJ.min(axis=(1, 2))
    min_delta = np.minimum(min_h, min_j)
    out = np.empty((h.shape[0], 2), dtype=np.float64)
    empty = ~np.isfinite(min_delta)
    with np.errstate(divide="ignore"):
        out[:, 0] = np.log(2) / max_delta
        out[:, 1] = np.log(100) / min_delta
    out[empty] = (0.1, 1.0)
    return out


def beta_range_from_reductions(min_delta: np.ndarray, max_delta: np.ndarray) -> np.ndarray:
    """The same rule from its two reductions (K0's ``range`` output: smallest non-zero |bias|, largest total |bias|;
    both 0 when every bias is zero)."""
    mn = np.atleast_1d(np.asarray(min_delta, dtype=np.float64))
    mx = np.atleast_1d(np.asarray(max_delta, dtype=np.float64))
    out = np.empty((mn.shape[0], 2), dtype=np.float64)
    empty = mx == 0.0
    with np.errstate(divide="ignore"):
        out[:, 0] = np.log(2) / mx
        out[:, 1] = np.log(100) / mn
    out[empty] = (0.1, 1.0)
    return out


def beta_schedule(beta_range: np.ndarray, num_sweeps: int, beta_schedule_type: str = "geometric"):
    """Returns ``(betas [B, num_betas] f64, sweeps_per_beta)`` exactly as neal's ``sample()`` builds them."""
    beta_range = np.atleast_2d(np.asarray(beta_range, dtype=np.float64))
    if num_sweeps < 0:
        raise ValueError("num_sweeps should be non-negative")
    sweeps_per_beta = int(max(1, num_sweeps // 1000.0))
    num_betas = int(math.ceil(num_sweeps / sweeps_per_beta))
    if num_betas == 0:
        return np.zeros((beta_range.shape[0], 0)), sweeps_per_beta
    if beta_schedule_type == "geometric":
        betas = np.geomspace(beta_range[:, 0], beta_range[:, 1], num_betas, axis=1)
    elif beta_schedule_type == "linear":
        betas = np.linspace(beta_range[:, 0], beta_range[:, 1], num_betas, axis=1)
    else:
        raise ValueError("Beta schedule type {} not implemented".format(beta_schedule_type))
    return np.ascontiguousarray(betas), sweeps_per_beta


def initial_states_numpy(seed, num_reads: int, n: int) -> np.ndarray:
    """dimod's random generator: int8 [num_reads, n] of 0/1 (0 = spin -1)."""
    rs = np.random.RandomState(seed)
    s = rs.choice(sorted([-1, 1]), size=(num_reads, n))
    return (s > 0).astype(np.int8)


def check_seed(seed):
    """neal/dimod accept None or an int in [0, 2**32) (numpy RandomState's range)."""
    if seed is None:
        return None
    if isinstance(seed, (bool, np.bool_)) or not isinstance(seed, (int, np.integer)):
        raise TypeError("'seed' should be None or a positive integer")
    if not (0 <= int(seed) <= 2 ** 32 - 1):
        raise ValueError("'seed' should be an integer between 0 and 2^32 - 1")
    return int(seed)
