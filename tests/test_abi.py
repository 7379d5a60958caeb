"""CPU: the C-ABI shared library loads and exports every symbol include/qbm_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qbm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qbm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("qbm_sa_sample", "qbm_qubo_energy", "qbm_phase_stats", "qbm_qubo_to_ising", "qbm_last_error", "qbm_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(qbm):
    L = qbm._lib.load()
    for s in declared_symbols():
        assert hasattr(L, s), f"libqbm_b200.so does not export {s}"
        assert s in qbm._lib.SIGNATURES, f"_lib.py has no ctypes signature for {s}"
    assert L.qbm_version() == 1


def test_argument_errors_do_not_need_a_gpu(qbm):
    """Validation happens before any CUDA call: error codes and messages cross the ABI intact."""
    import pytest
    L = qbm._lib.load()
    rc = L.qbm_sa_sample(None, None, 4, 4, 1, None, 0, 1, 1, 1, 0, 0, None, None, None, None, 0, 0, None)
    assert rc == qbm._lib.QBM_EINVAL
    assert b"null pointer" in L.qbm_last_error()
    with pytest.raises(ValueError):
        qbm._lib.check(rc)
    assert L.qbm_sa_workspace_bytes(2048, 1) == (2048 + 1) * 2048 * 4
    assert L.qbm_sa_workspace_bytes(24, 3) == 3 * 25 * 128 * 4
    assert L.qbm_sa_workspace_bytes(0, 1) == 0
    # the two-phase workspace adds the hand-over buffers (fields + sweep counters) only where that schedule exists
    assert L.qbm_sa_workspace_bytes_two_phase(896, 2, 100) == L.qbm_sa_workspace_bytes(896, 2)          # 7 windows: not supported
    assert L.qbm_sa_workspace_bytes_two_phase(256, 2, 100) == L.qbm_sa_workspace_bytes(256, 2)
    assert L.qbm_sa_workspace_bytes_two_phase(512, 2, 100) == L.qbm_sa_workspace_bytes(512, 2) + 200 * 512 * 4 + 800
    assert L.qbm_sa_workspace_bytes_two_phase(1024, 2, 100) == L.qbm_sa_workspace_bytes(1024, 2) + 200 * 1024 * 4 + 800
    assert L.qbm_sa_workspace_bytes_two_phase(2048, 2, 100) == L.qbm_sa_workspace_bytes(2048, 2) + 200 * 2048 * 4 + 800
    assert L.qbm_sa_workspace_bytes_two_phase(1800, 1, 3) == L.qbm_sa_workspace_bytes(1800, 1) + 3 * 2048 * 4 + 16
    assert L.qbm_phase_stats_workspace_bytes(2, 100, 34) == 2 * 34 * 4 * 4
    dummy = ctypes.c_void_p(16)
    rc = L.qbm_sa_sample(dummy, dummy, 4096, 4096, 1, dummy, 0, 1, 1, 1, 0, 0, None, dummy, None, dummy, 1 << 40, 0, None)
    assert rc == qbm._lib.QBM_EUNSUPPORTED and b"QBM_SA_MAX_N" in L.qbm_last_error()
    rc = L.qbm_sa_sample(dummy, dummy, 64, 64, 1, dummy, 0, 1, 1, 1, 0, 0, None, dummy, None, dummy, 16, 0, None)
    assert rc == qbm._lib.QBM_EWORKSPACE
