"""Duck-typed ``dimod`` / ``neal`` modules (boundary B2) so that the reference's
``Disc_QBM`` (src/model/faster_dqbm.py, src/model/discriminative_qbm.py) and ``LocalSASampler``
(src/qubo/sampler.py) run unmodified on top of the B200 sampler.

``install()`` registers them in ``sys.modules`` under the names the reference imports (``dimod``,
``neal``).  It refuses to shadow a real installation unless ``force=True``.
"""
from __future__ import annotations

import importlib.util
import sys

from . import dimod_shim, neal_shim


def install(force: bool = False) -> None:
    for name, mod in (("dimod", dimod_shim), ("neal", neal_shim)):
        present = name in sys.modules and sys.modules[name] is not mod
        if not present and not force:
            try:
                present = importlib.util.find_spec(name) is not None
            except (ImportError, ValueError):
                present = False
        if present and not force:
            raise RuntimeError(f"a real '{name}' is importable; pass force=True to shadow it with the B200 shim")
        sys.modules[name] = mod


def uninstall() -> None:
    for name, mod in (("dimod", dimod_shim), ("neal", neal_shim)):
        if sys.modules.get(name) is mod:
            del sys.modules[name]
