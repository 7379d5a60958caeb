"""oracle/ref_stubs.py -- TEST INFRASTRUCTURE ONLY.

Lets the reference's own Python (``/root/reference/src/...``) be imported in THIS container, where
its third-party dependencies are absent, without editing it: throw-away stub modules are injected
into ``sys.modules`` (``dimod``, ``neal``, ``dwave.*``, ``dwave_networkx``, ``minorminer``,
``pymetis``, ``matplotlib``, ``seaborn``).  The ``neal`` stub's sampler is the oracle's C
restatement of dwave-neal 0.5.9 (oracle/neal_sa.c), so reference code + these stubs = the CPU
reference stack used to generate tests/golden/ and to time the reference's training loop.

Nothing here is reachable from the product package.  /root/reference only exists in the build
container: everything that needs it is guarded by ``reference_available()``.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from collections.abc import Mapping

import numpy as np

REFERENCE_ROOT = os.environ.get("QBM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


class _View(Mapping):
    def __init__(self, row):
        self._r = row

    def __getitem__(self, k):
        return self._r[k]

    def __iter__(self):
        return iter(range(len(self._r)))

    def __len__(self):
        return len(self._r)

    def values(self):
        return self._r.tolist()


class StubBQM:
    """dimod.BQM(Q, "BINARY") for a dense ndarray."""

    def __init__(self, Q, vartype):
        assert str(vartype).upper().endswith("BINARY")
        self.Q = np.asarray(Q, dtype=np.float64)
        self.vartype = "BINARY"
        n = self.Q.shape[0]
        B = self.Q + self.Q.T
        self.linear = {i: float(self.Q[i, i]) for i in range(n)}
        self.quadratic = {(i, j): float(B[i, j]) for i in range(n) for j in range(i + 1, n) if B[i, j] != 0.0}
        self.num_variables = n

    def energies(self, X):
        X = np.asarray(X, dtype=np.float64)
        U = np.triu(self.Q + self.Q.T, k=1) + np.diag(np.diag(self.Q))
        return np.einsum("ri,ij,rj->r", X, U, X)


class StubSampleSet:
    def __init__(self, samples, energies):
        R, n = samples.shape
        self.record = np.rec.fromarrays(
            [samples.astype(np.int8), np.asarray(energies, dtype=np.float64), np.ones(R, dtype=np.intc)],
            dtype=[("sample", np.int8, (n,)), ("energy", np.float64), ("num_occurrences", np.intc)])
        self.variables = list(range(n))

    @classmethod
    def from_samples_bqm(cls, samples_like, bqm):
        rows = np.array([[s[v] for v in range(bqm.num_variables)] if isinstance(s, Mapping) else list(s)
                         for s in samples_like])
        return cls(rows, bqm.energies(rows))

    def samples(self):
        order = np.argsort(self.record.energy, kind="stable")
        return [_View(r) for r in self.record.sample[order]]

    def __len__(self):
        return self.record.shape[0]


class StubNealSampler:
    """neal.SimulatedAnnealingSampler backed by the oracle's cpu_sa.cpp restatement."""

    def sample(self, bqm, num_reads=None, num_sweeps=1000, seed=None, beta_range=None, **_):
        from . import oracle as O
        s, e = O.neal_sample(bqm.Q, int(num_reads or 1), int(num_sweeps), seed=seed, beta_range=beta_range)
        return StubSampleSet(s, e)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def stub_modules() -> dict:
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Anything()

    anything = _Anything()

    def lazy(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return anything
    dimod = _mod("dimod", BQM=StubBQM, BinaryQuadraticModel=StubBQM, SampleSet=StubSampleSet, BINARY="BINARY", SPIN="SPIN")
    neal = _mod("neal", SimulatedAnnealingSampler=StubNealSampler)
    plt = _mod("matplotlib.pyplot")
    plt.__getattr__ = lazy  # type: ignore[attr-defined]
    mpl = _mod("matplotlib", pyplot=plt)
    mpl.__getattr__ = lazy  # type: ignore[attr-defined]
    dwave = _mod("dwave")
    cloud = _mod("dwave.cloud", Client=_Anything)
    emb = _mod("dwave.embedding", embed_bqm=anything, unembed_sampleset=anything, EmbeddedStructure=_Anything)
    dwave.cloud, dwave.embedding = cloud, emb
    mods = {"dimod": dimod, "neal": neal, "matplotlib": mpl, "matplotlib.pyplot": plt, "dwave": dwave,
            "dwave.cloud": cloud, "dwave.embedding": emb}
    for name in ("dwave_networkx", "minorminer", "pymetis", "seaborn"):
        m = _mod(name)
        m.__getattr__ = lazy  # type: ignore[attr-defined]
        mods[name] = m
    return mods


@contextlib.contextmanager
def reference_imports(extra: dict | None = None):
    """Context manager: stubs + /root/reference on sys.path + CWD holding src/secrets/TOKEN.txt
    (``Disc_QBM.__init__`` opens that relative path, faster_dqbm.py:72)."""
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (it only exists in the build container)")
    import torch  # noqa: F401  (real module; must be imported before the stubs go in)
    import sklearn.metrics  # noqa: F401
    mods = stub_modules()
    mods.update(extra or {})
    saved = {k: sys.modules.get(k) for k in mods}
    saved_src = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved_src:
        del sys.modules[k]
    sys.modules.update(mods)
    sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)          # src/secrets/TOKEN.txt (empty) lives there
    try:
        yield
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved_src)
