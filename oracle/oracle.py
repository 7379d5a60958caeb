"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's sampling path, used as the checker by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs.  Nothing in the product
package (``qbm-image-classification_b200/``) imports this module.

What is restated, and from where (citations into /root/reference and SURVEY.md Appendix A --
``dimod==0.12.18`` / ``dwave-neal==0.5.9`` are pinned in requirements.txt:1,3 but are not
vendored and not installable here, so their published algorithms are restated):

* ``qubo_to_ising``        dimod.BQM(Q, "BINARY") + change_vartype(SPIN)  (A.1, A.2;
                            call sites src/qubo/sampler.py:7-8, src/model/faster_dqbm.py:577,619)
* ``default_beta_range``   neal/sampler.py:281 ``_default_ising_beta_range`` LEGACY rule (A.4)
* ``beta_schedule``        neal ``sample()`` argument handling (A.3)
* ``initial_states``       dimod/core/initialized.py:207 ``_random_generator`` (A.3)
* ``neal_sample``          the whole ``SimulatedAnnealingSampler.sample`` call (A.3-A.6) on top of
                            ``oracle_neal_sa`` (neal_sa.c = cpu_sa.cpp restated)
* ``replay_sample``        the replay oracle (replay_sa.c): reference Metropolis rule in the
                            kernel's fp32 arithmetic and Philox stream -> bit-exact target
* ``sample_Q_reference``   src/qubo/sampler.py:26-33 ``LocalSASampler.sample_Q`` incl. the
                            linear-only shortcut (:13-17)

PARITY STATUS: sample-level behaviour of neal is "parity unpinned" (no golden sample sets in the
reference, neal absent); ``neal_sample`` is pinned end-to-end by reference-held data -- the 70
PneumoniaMNIST last-epoch runs (weights + recorded accuracy/AUC) are annealed through it and every
recorded pair is reproduced exactly (tests/test_oracle_models.py) -- and by exact identities
(tests/test_oracle_sa.py).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with gcc (idempotent)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("neal_sa.c", "replay_sa.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c_i, c_u64, c_p = ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p
        L.oracle_neal_sa.argtypes = [c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_u64, c_p, c_p]
        L.oracle_neal_sa.restype = c_i
        L.oracle_replay_sa.argtypes = [c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_u64, c_u64, c_i, c_p, c_p, c_p]
        L.oracle_replay_sa.restype = c_i
        L.oracle_philox4x32_10.argtypes = [c_p, c_p, c_p]
        L.oracle_philox4x32_10.restype = None
        L.oracle_neg_log_u32.argtypes = [ctypes.c_uint32]
        L.oracle_neg_log_u32.restype = ctypes.c_float
        L.oracle_qubo_energy.argtypes = [c_i, c_p, ctypes.c_longlong, c_p, c_p]
        L.oracle_qubo_energy.restype = c_i
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------
# dimod glue (A.1, A.2)
# --------------------------------------------------------------------------------------------
def qubo_to_ising(Q: np.ndarray):
    """Dense QUBO -> spin model, float64.

    Returns ``(h[n], Jsym[n,n], offset, irow, icol, jval)``: ``Jsym`` is the symmetric coupling
    matrix with zero diagonal (``Jsym[i,j] = J_ij``); the triplets list couplers with i<j and a
    non-zero bias only, as dimod stores them (A.1).  Energies: ``x^T Q x = h.s + sum_{i<j} J_ij
    s_i s_j + offset`` for ``x = (s+1)/2``.
    """
    Q = np.asarray(Q, dtype=np.float64)
    n = Q.shape[0]
    assert Q.shape == (n, n)
    a = np.diag(Q).copy()
    B = Q + Q.T
    np.fill_diagonal(B, 0.0)               # b_ij = Q_ij + Q_ji for i != j
    Jsym = B / 4.0
    h = a / 2.0 + B.sum(axis=1) / 4.0
    iu = np.triu_indices(n, k=1)
    b_up = B[iu]
    offset = a.sum() / 2.0 + b_up.sum() / 4.0
    nz = b_up != 0.0
    irow = iu[0][nz].astype(np.int32)
    icol = iu[1][nz].astype(np.int32)
    jval = (b_up[nz] / 4.0).astype(np.float64)
    return h, Jsym, float(offset), irow, icol, jval


def default_beta_range(h: np.ndarray, jval: np.ndarray, irow: np.ndarray, icol: np.ndarray):
    """neal 0.5.9 ``_default_ising_beta_range`` (legacy rule, A.4) on SPIN biases."""
    abs_h = np.abs(h[h != 0])
    abs_j = np.abs(jval[jval != 0])
    if abs_h.size + abs_j.size == 0:
        return [0.1, 1.0]
    min_delta = min(abs_h.min() if abs_h.size else math.inf, abs_j.min() if abs_j.size else math.inf)
    tot = np.abs(h).astype(np.float64).copy()
    np.add.at(tot, irow, np.abs(jval))
    np.add.at(tot, icol, np.abs(jval))
    max_delta = tot.max()
    return [float(np.log(2) / max_delta), float(np.log(100) / min_delta)]


def beta_schedule(beta_range, num_sweeps: int):
    """neal ``sample()``: sweeps-per-beta rule and geometric schedule (A.3)."""
    sweeps_per_beta = int(max(1, num_sweeps // 1000.0))
    num_betas = int(math.ceil(num_sweeps / sweeps_per_beta))
    betas = np.geomspace(beta_range[0], beta_range[1], num_betas)
    return betas.astype(np.float64), sweeps_per_beta


def initial_states(seed, num_reads: int, n: int) -> np.ndarray:
    """dimod ``_random_generator``: RandomState(seed).choice([-1, 1], size=(reads, n)) int8 (A.3)."""
    rs = np.random.RandomState(seed)
    return rs.choice(sorted([-1, 1]), size=(num_reads, n)).astype(np.int8)


# --------------------------------------------------------------------------------------------
# neal restatement
# --------------------------------------------------------------------------------------------
def neal_sample(Q: np.ndarray, num_reads: int, num_sweeps: int = 1000, seed=None, beta_range=None,
                return_info: bool = False):
    """``neal.SimulatedAnnealingSampler().sample(dimod.BQM(Q, "BINARY"), num_reads=, num_sweeps=,
    seed=)`` restated: returns ``(samples int8 [R,n] of 0/1 in read order, energies f64 [R])``."""
    Q = np.asarray(Q, dtype=np.float64)
    n = Q.shape[0]
    h, _, offset, irow, icol, jval = qubo_to_ising(Q)
    if beta_range is None:
        beta_range = default_beta_range(h, jval, irow, icol)
    betas, spb = beta_schedule(beta_range, num_sweeps)
    if seed is None:
        seed = int(np.random.randint(2 ** 31))
    states = np.ascontiguousarray(initial_states(seed, num_reads, n))
    energies = np.zeros(num_reads, dtype=np.float64)
    counters = np.zeros(3, dtype=np.uint64)
    h = np.ascontiguousarray(h)
    rc = lib().oracle_neal_sa(n, _ptr(h), int(len(jval)), _ptr(irow), _ptr(icol), _ptr(jval),
                              int(num_reads), _ptr(states), int(len(betas)), _ptr(betas), int(spb),
                              int(seed), _ptr(energies), _ptr(counters))
    if rc:
        raise RuntimeError(f"oracle_neal_sa failed: {rc}")
    samples = ((states + 1) // 2).astype(np.int8)
    energies = energies + offset
    if return_info:
        return samples, energies, {"beta_range": list(beta_range), "counters": counters,
                                   "num_betas": len(betas), "sweeps_per_beta": spb}
    return samples, energies


def sample_Q_reference(Q: np.ndarray, num_reads: int, num_sweeps: int = 1000, seed=None) -> np.ndarray:
    """src/qubo/sampler.py:26-33 restated (float32 [R,n] of 0/1, read order), including the
    linear-only shortcut :13-17 (ground state replicated; rng coin for zero biases)."""
    Q = np.asarray(Q, dtype=np.float64)
    n = Q.shape[0]
    B = Q + Q.T
    np.fill_diagonal(B, 0.0)
    if not np.any(B != 0.0):
        rng = np.random.default_rng(seed)
        sol = np.zeros(n, dtype=np.float32)
        for v in range(n):                      # dict-comprehension order = variable order
            hv = Q[v, v]
            sol[v] = 1 if hv < 0 else (0 if hv > 0 else int(rng.integers(0, 2)))
        return np.tile(sol, (int(num_reads), 1)).astype(np.float32)
    s, _ = neal_sample(Q, num_reads, num_sweeps, seed)
    return s.astype(np.float32)


# --------------------------------------------------------------------------------------------
# replay oracle
# --------------------------------------------------------------------------------------------
def replay_sample(J32: np.ndarray, h32: np.ndarray, betas32: np.ndarray, sweeps_per_beta: int, seed: int,
                  chain_first: int, num_chains: int, init01: np.ndarray | None = None, n: int | None = None):
    """Replay oracle on one problem.  ``J32`` is [n, ld] fp32 (rows may be zero-padded), returns
    ``(states int8 [num_chains, n] of 0/1, counters u64[3] = accepted, skipped, draws)``."""
    J32 = np.ascontiguousarray(J32, dtype=np.float32)
    h32 = np.ascontiguousarray(h32, dtype=np.float32)
    betas32 = np.ascontiguousarray(betas32, dtype=np.float32)
    if n is None:
        n = J32.shape[0]
    ld = J32.shape[1]
    out = np.zeros((num_chains, n), dtype=np.int8)
    counters = np.zeros(3, dtype=np.uint64)
    if init01 is not None:
        init01 = np.ascontiguousarray(init01, dtype=np.int8)
        assert init01.shape == (num_chains, n)
    rc = lib().oracle_replay_sa(int(n), int(ld), _ptr(J32), _ptr(h32), int(len(betas32)), _ptr(betas32),
                                int(sweeps_per_beta), int(seed), int(chain_first), int(num_chains),
                                _ptr(init01), _ptr(out), _ptr(counters))
    if rc:
        raise RuntimeError(f"oracle_replay_sa failed: {rc}")
    return out, counters


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return o


def neg_log_u32(u: int) -> float:
    return float(lib().oracle_neg_log_u32(ctypes.c_uint32(int(u))))


def qubo_energies(Q: np.ndarray, X01: np.ndarray) -> np.ndarray:
    """x^T Q x for every row of X01 (float64) -- the energy dimod reports for BINARY samples (A.2)."""
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    X = np.ascontiguousarray(X01, dtype=np.int8)
    out = np.zeros(X.shape[0], dtype=np.float64)
    lib().oracle_qubo_energy(int(Q.shape[0]), _ptr(Q), int(X.shape[0]), _ptr(X), _ptr(out))
    return out
