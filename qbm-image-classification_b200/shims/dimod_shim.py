"""Minimal duck types of the ``dimod`` objects the reference touches (boundary types only -- this is
not a re-implementation of dimod):

* ``BQM(Q, "BINARY")`` from a dense ndarray (src/qubo/sampler.py:8; src/model/faster_dqbm.py:577,619):
  ``linear[i] = Q[i,i]``, ``quadratic[i,j] = Q[i,j] + Q[j,i]`` stored only when non-zero
  (SURVEY.md Appendix A.1), so ``len(bqm.quadratic) == 0`` / ``bqm.quadratic == {}`` hold for
  diagonal QUBOs (sampler.py:10-11, faster_dqbm.py:43).
* ``SampleSet``: ``.samples()`` (sized iterable of mappings, energy-sorted by default, ``.values()`` in
  variable order -- faster_dqbm.py:300-301,697-698,777-778), ``.record.sample / .energy /
  .num_occurrences`` in read order (sampler.py:33), ``.variables``, ``.vartype``, ``.info``, ``.first``,
  ``from_samples`` and ``from_samples_bqm`` (sampler.py:17, faster_dqbm.py:53).
"""
from __future__ import annotations

import enum
from collections import namedtuple
from collections.abc import Mapping

import numpy as np

__version__ = "0.12.18+qbm_b200.shim"


class Vartype(enum.Enum):
    SPIN = frozenset({-1, 1})
    BINARY = frozenset({0, 1})


SPIN = Vartype.SPIN
BINARY = Vartype.BINARY


def as_vartype(v) -> Vartype:
    if isinstance(v, Vartype):
        return v
    if isinstance(v, str):
        try:
            return Vartype[v.upper()]
        except KeyError:
            pass
    if isinstance(v, (set, frozenset)):
        for vt in Vartype:
            if frozenset(v) == vt.value:
                return vt
    raise TypeError("expected input vartype to be one of: Vartype.SPIN, 'SPIN', {-1, 1}, Vartype.BINARY, 'BINARY', or {0, 1}.")


class BinaryQuadraticModel:
    """Dense-backed binary quadratic model over variables ``0..n-1``."""

    def __init__(self, *args, vartype=None, offset=0.0):
        if len(args) == 2 and not isinstance(args[0], Mapping):
            Q, vartype = args
            Q = np.asarray(Q, dtype=np.float64)
            if Q.ndim != 2 or Q.shape[0] != Q.shape[1]:
                raise ValueError("expected dense to be a 2 dim square array")
            self._lin = np.diag(Q).astype(np.float64).copy()
            B = Q + Q.T
            np.fill_diagonal(B, 0.0)
            self._quad = B            # symmetric, zero diagonal: b_ij
        elif len(args) == 3 or (len(args) == 4):
            linear, quadratic = args[0], args[1]
            if len(args) == 4:
                offset = args[2]
            vartype = args[-1]
            labels = sorted(set(linear) | {u for u, _ in quadratic} | {v for _, v in quadratic})
            if labels != list(range(len(labels))):
                raise ValueError("this shim supports integer variable labels 0..n-1 only")
            n = len(labels)
            self._lin = np.zeros(n)
            for v, b in linear.items():
                self._lin[v] += b
            self._quad = np.zeros((n, n))
            for (u, v), b in quadratic.items():
                if u == v:
                    raise ValueError("self-loops are not allowed in a quadratic bias")
                self._quad[u, v] += b
                self._quad[v, u] += b
        else:
            raise TypeError("unsupported BQM constructor arguments for the qbm_b200 dimod shim")
        self.vartype = as_vartype(vartype)
        self.offset = float(offset)

    # -- views -------------------------------------------------------------------------------
    @property
    def num_variables(self) -> int:
        return int(self._lin.shape[0])

    def __len__(self) -> int:
        return self.num_variables

    @property
    def variables(self):
        return range(self.num_variables)

    @property
    def linear(self) -> dict:
        return {i: float(b) for i, b in enumerate(self._lin)}

    @property
    def quadratic(self) -> dict:
        iu, ju = np.nonzero(np.triu(self._quad, k=1))
        return {(int(i), int(j)): float(self._quad[i, j]) for i, j in zip(iu, ju)}

    @property
    def num_interactions(self) -> int:
        return int(np.count_nonzero(np.triu(self._quad, k=1)))

    @property
    def shape(self):
        return self.num_variables, self.num_interactions

    def to_qubo_matrix(self) -> np.ndarray:
        """Upper-triangular dense QUBO (BINARY models)."""
        if self.vartype is not Vartype.BINARY:
            return self.change_vartype(Vartype.BINARY, inplace=False).to_qubo_matrix()
        return np.triu(self._quad, k=1) + np.diag(self._lin)

    def change_vartype(self, vartype, inplace: bool = True):
        vartype = as_vartype(vartype)
        tgt = self if inplace else self.copy()
        if vartype is tgt.vartype:
            return tgt
        lin, quad, off = tgt._lin, tgt._quad, tgt.offset
        up = np.triu(quad, k=1)
        if vartype is Vartype.SPIN:          # x = (s + 1) / 2
            tgt._lin = lin / 2.0 + quad.sum(axis=1) / 4.0
            tgt._quad = quad / 4.0
            tgt.offset = off + lin.sum() / 2.0 + up.sum() / 4.0
        else:                                # s = 2x - 1
            tgt._lin = 2.0 * lin - 2.0 * quad.sum(axis=1)
            tgt._quad = 4.0 * quad
            tgt.offset = off - lin.sum() + up.sum()
        tgt.vartype = vartype
        return tgt

    def copy(self):
        new = object.__new__(BinaryQuadraticModel)
        new._lin = self._lin.copy()
        new._quad = self._quad.copy()
        new.vartype = self.vartype
        new.offset = self.offset
        return new

    def energies(self, samples) -> np.ndarray:
        S = _samples_array(samples, self.num_variables).astype(np.float64)
        return S @ self._lin + np.einsum("ri,ij,rj->r", S, np.triu(self._quad, k=1), S) + self.offset

    def energy(self, sample) -> float:
        return float(self.energies([sample])[0])


BQM = BinaryQuadraticModel


def _samples_array(samples, n: int) -> np.ndarray:
    if isinstance(samples, tuple) and len(samples) == 2:
        samples = samples[0]
    if isinstance(samples, np.ndarray):
        arr = samples
    else:
        samples = list(samples)
        if samples and isinstance(samples[0], Mapping):
            arr = np.array([[s[v] for v in range(n)] for s in samples])
        else:
            arr = np.asarray(samples)
    arr = np.atleast_2d(arr)
    if arr.size == 0:
        arr = arr.reshape(0, n)
    return arr


class SampleView(Mapping):
    """One row of a SampleSet as a read-only mapping variable -> value (variable order)."""
    __slots__ = ("_row",)

    def __init__(self, row):
        self._row = row

    def __getitem__(self, v):
        return self._row[v]

    def __iter__(self):
        return iter(range(len(self._row)))

    def __len__(self):
        return len(self._row)

    def values(self):
        return self._row.tolist()

    def __repr__(self):
        return repr(dict(self.items()))


class SamplesArray:
    """What ``SampleSet.samples()`` returns: sized, iterable, indexable collection of mappings."""

    def __init__(self, rows: np.ndarray):
        self._rows = rows

    def __len__(self):
        return self._rows.shape[0]

    def __iter__(self):
        return (SampleView(r) for r in self._rows)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return SampleView(self._rows[idx])
        return self._rows[idx]


Sample = namedtuple("Sample", ["sample", "energy", "num_occurrences"])


class SampleSet:
    def __init__(self, record: np.recarray, variables, info: dict, vartype: Vartype):
        self.record = record
        self.variables = list(variables)
        self.info = info
        self.vartype = vartype

    @classmethod
    def from_samples(cls, samples_like, energy, vartype, info=None, num_occurrences=None, **_):
        labels = None
        if isinstance(samples_like, tuple) and len(samples_like) == 2:
            samples_like, labels = samples_like
        if isinstance(samples_like, np.ndarray):
            arr = np.atleast_2d(samples_like)
        else:
            samples_like = list(samples_like)
            if samples_like and isinstance(samples_like[0], Mapping):
                keys = sorted(samples_like[0])
                labels = keys if labels is None else labels
                arr = np.array([[s[k] for k in keys] for s in samples_like])
            else:
                arr = np.atleast_2d(np.asarray(samples_like))
        R, n = arr.shape
        energy = np.broadcast_to(np.asarray(energy, dtype=np.float64), (R,))
        occ = np.ones(R, dtype=np.intc) if num_occurrences is None else np.asarray(num_occurrences, dtype=np.intc)
        rec = np.rec.fromarrays([arr.astype(np.int8), energy, occ],
                                dtype=[("sample", np.int8, (n,)), ("energy", np.float64), ("num_occurrences", np.intc)])
        return cls(rec, list(range(n)) if labels is None else list(labels), dict(info or {}), as_vartype(vartype))

    @classmethod
    def from_samples_bqm(cls, samples_like, bqm, **kwargs):
        arr = _samples_array(samples_like, bqm.num_variables)
        return cls.from_samples(arr, bqm.energies(arr), bqm.vartype, **kwargs)

    def __len__(self):
        return int(self.record.shape[0])

    def samples(self, n=None, sorted_by="energy"):
        rows = self.record.sample
        if sorted_by is not None:
            order = np.argsort(self.record[sorted_by], kind="stable")
            rows = rows[order]
        if n is not None:
            rows = rows[:n]
        return SamplesArray(rows)

    def data(self, fields=None, sorted_by="energy", reverse=False, **_):
        order = np.arange(len(self)) if sorted_by is None else np.argsort(self.record[sorted_by], kind="stable")
        if reverse:
            order = order[::-1]
        for i in order:
            yield Sample(SampleView(self.record.sample[i]), float(self.record.energy[i]),
                         int(self.record.num_occurrences[i]))

    def __iter__(self):
        return iter(self.samples())

    @property
    def first(self):
        return next(self.data())

    def lowest(self, rtol=1.e-5, atol=1.e-8):
        e = self.record.energy
        keep = np.isclose(e, e.min(), rtol=rtol, atol=atol) if len(e) else np.zeros(0, bool)
        return SampleSet(self.record[keep], self.variables, dict(self.info), self.vartype)

    def aggregate(self):
        rows, first_idx, inv = np.unique(self.record.sample, axis=0, return_index=True, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        occ = np.bincount(inv, weights=self.record.num_occurrences, minlength=rows.shape[0]).astype(np.intc)
        out = SampleSet.from_samples(rows, self.record.energy[first_idx], self.vartype, info=self.info,
                                     num_occurrences=occ)
        out.variables = list(self.variables)
        return out

    def change_vartype(self, vartype, energy_offset=0.0, inplace=True):
        vartype = as_vartype(vartype)
        tgt = self if inplace else SampleSet(self.record.copy(), self.variables, dict(self.info), self.vartype)
        if vartype is tgt.vartype:
            return tgt
        if vartype is Vartype.BINARY:
            tgt.record.sample[:] = (tgt.record.sample + 1) // 2
        else:
            tgt.record.sample[:] = 2 * tgt.record.sample - 1
        tgt.record.energy[:] = tgt.record.energy + energy_offset
        tgt.vartype = vartype
        return tgt
