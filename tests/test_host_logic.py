"""CPU: the product's host-side logic (ising.py, the dimod/neal shims, the sampler's host branches)
against the oracle's independent restatement."""
import pickle

import numpy as np
import pytest

from conftest import random_qubo


def test_qubo_to_ising_matches_oracle(qbm, oracle):
    rng = np.random.default_rng(0)
    for n in (1, 4, 24, 193):
        Q = rng.uniform(-1, 1, (n, n)) if n == 4 else random_qubo(n, seed=n, density=0.8)
        h, J, off = qbm.ising.qubo_to_ising(Q)
        ho, Jo, offo, irow, icol, jv = oracle.qubo_to_ising(Q)
        assert np.array_equal(J[0], Jo) and np.allclose(h[0], ho, rtol=1e-15, atol=1e-15)
        assert np.isclose(off[0], offo, rtol=1e-13, atol=1e-13)
        br = qbm.ising.default_beta_range(h, J)[0]
        bro = oracle.default_beta_range(ho, jv, irow, icol)
        assert np.allclose(br, bro, rtol=1e-14)


def test_batched_schedule_equals_scalar_geomspace(qbm, oracle):
    Qs = np.stack([random_qubo(30, seed=s, scale=1 + s) for s in range(4)])
    h, J, _ = qbm.ising.qubo_to_ising(Qs)
    br = qbm.ising.default_beta_range(h, J)
    betas, spb = qbm.ising.beta_schedule(br, 1000)
    assert betas.shape == (4, 1000) and spb == 1
    for b in range(4):
        ref, spbo = oracle.beta_schedule(br[b], 1000)
        assert np.array_equal(betas[b], ref) and spbo == spb
    assert qbm.ising.beta_schedule(br, 20)[0].shape == (4, 20)
    assert qbm.ising.beta_schedule(br, 2500) [1] == 2
    lin, _ = qbm.ising.beta_schedule(br[:1], 10, "linear")
    assert np.allclose(lin[0], np.linspace(br[0, 0], br[0, 1], 10))
    with pytest.raises(ValueError):
        qbm.ising.beta_schedule(br, 10, "bogus")
    z = qbm.ising.default_beta_range(np.zeros((1, 3)), np.zeros((1, 3, 3)))
    assert z.tolist() == [[0.1, 1.0]]


def test_initial_states_are_dimods(qbm, oracle):
    a = qbm.ising.initial_states_numpy(44, 7, 13)
    b = oracle.initial_states(44, 7, 13)
    assert np.array_equal(2 * a.astype(int) - 1, b)


def test_seed_validation_like_neal(qbm):
    assert qbm.ising.check_seed(None) is None and qbm.ising.check_seed(5) == 5
    for bad in (-1, 2 ** 32):
        with pytest.raises(ValueError):
            qbm.ising.check_seed(bad)
    for bad in (1.5, "3", True):
        with pytest.raises(TypeError):
            qbm.ising.check_seed(bad)


def test_linear_only_shortcut_is_a_host_branch(qbm, oracle):
    """src/qubo/sampler.py:13-17,28-29: diagonal QUBOs never reach the annealer (no GPU needed)."""
    s = qbm.B200SASampler(num_sweeps=1000, seed=44)
    Q = np.diag([1.0, -2.0, 0.0, 3.0, 0.0])
    out = s.sample_Q(Q, 6)
    assert out.dtype == np.float32 and out.shape == (6, 5)
    assert np.array_equal(out, oracle.sample_Q_reference(Q, 6, 1000, seed=44))
    with pytest.raises(ValueError):
        s.sample_Q(np.zeros((3, 4)), 2)


def test_dimod_shim_boundary_types(qbm):
    from qbm_b200.shims import dimod_shim as dimod
    Q = np.triu(np.arange(1.0, 10.0).reshape(3, 3))
    b = dimod.BQM(Q, "BINARY")
    assert b.linear == {0: 1.0, 1: 5.0, 2: 9.0} and b.quadratic == {(0, 1): 2.0, (0, 2): 3.0, (1, 2): 6.0}
    assert len(dimod.BQM(np.diag([1.0, 2.0]), "BINARY").quadratic) == 0
    assert dimod.BQM(np.diag([1.0, 2.0]), "BINARY").quadratic == {}
    X = np.array([[1, 0, 1], [1, 1, 1], [0, 0, 0]])
    e = b.energies(X)
    assert np.allclose(e, np.einsum("ri,ij,rj->r", X, Q, X))
    sp = b.change_vartype("SPIN", inplace=False)
    assert np.allclose(sp.energies(2 * X - 1), e)
    assert np.allclose(sp.change_vartype("BINARY", inplace=False).to_qubo_matrix(), Q)
    ss = dimod.SampleSet.from_samples_bqm([{0: 1, 1: 0, 2: 1}, {0: 0, 1: 0, 2: 0}, {0: 1, 1: 1, 2: 1}], b)
    assert ss.record.sample.tolist() == [[1, 0, 1], [0, 0, 0], [1, 1, 1]]          # read order
    assert [s.values() for s in ss.samples()] == [[0, 0, 0], [1, 0, 1], [1, 1, 1]]  # energy order
    assert len(ss.samples()) == 3 and ss.first.energy == 0.0 and ss.variables == [0, 1, 2]
    assert ss.record.num_occurrences.tolist() == [1, 1, 1]
    agg = dimod.SampleSet.from_samples(np.array([[1, 0], [1, 0], [0, 1]]), [1.0, 1.0, 2.0], "BINARY").aggregate()
    assert sorted(agg.record.num_occurrences.tolist()) == [1, 2]
    big = np.random.default_rng(3).integers(0, 2, (5000, 6)).astype(np.int8)      # 64 distinct rows, many repeats
    agg = dimod.SampleSet.from_samples(big, big.sum(1).astype(float), "BINARY", num_occurrences=np.full(5000, 2)).aggregate()
    assert len(agg) == len(np.unique(big, axis=0)) and int(agg.record.num_occurrences.sum()) == 10000
    assert np.array_equal(agg.record.energy, agg.record.sample.sum(1))
    low = ss.lowest()
    assert len(low) == 1 and low.record.sample.tolist() == [[0, 0, 0]]
    assert [d.energy for d in ss.data()] == sorted(ss.record.energy.tolist())
    assert [d.energy for d in ss.data(sorted_by=None)] == ss.record.energy.tolist()
    np_rows = np.vstack([np.array(list(s.values())) for s in ss.samples()])       # faster_dqbm.py:777-778
    assert np_rows.shape == (3, 3)
    pickle.loads(pickle.dumps(b))


def test_shim_install_and_neal_argument_checks(qbm):
    import sys
    qbm.shims.install()
    try:
        import dimod
        import neal
        assert dimod.BQM is qbm.shims.dimod_shim.BinaryQuadraticModel
        s = neal.SimulatedAnnealingSampler()
        pickle.loads(pickle.dumps(s))
        b = dimod.BQM(np.triu(np.ones((3, 3))), "BINARY")
        with pytest.raises(TypeError):
            s.sample(b, num_reads=2, num_sweeps=10, num_sweeps_per_beta=2)     # not in the 0.5.9 signature
        with pytest.raises(ValueError):
            s.sample(b, num_reads=2, num_sweeps=10, seed=-3)
        with pytest.raises(TypeError):
            s.sample(b, num_reads=2, num_sweeps=2.5)
    finally:
        qbm.shims.uninstall()
    assert "dimod" not in sys.modules and "neal" not in sys.modules


def test_destroy_process_group_releases_captured_data_parallel_steps():
    """NCCL cannot destroy a communicator while CUDA graphs that captured its collectives are alive: models with captured
    data-parallel steps register themselves and torch.distributed.destroy_process_group is wrapped to release them first."""
    import socket
    import torch.distributed as dist
    import qbm_b200.rbm as R

    class Fake:
        released = 0
        _graphs = {}

        def _close_peer(self, collective=True):
            assert collective is False          # the wrapper has already met the other ranks once
            self.released += 1

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        m = Fake()
        R._DP_GRAPH_MODELS.add(m)
        R._hook_destroy_process_group()
        R._hook_destroy_process_group()          # idempotent
    finally:
        dist.destroy_process_group()
    assert m.released == 1 and not dist.is_initialized()


def test_large_results_are_converted_to_float32_in_row_blocks(qbm):
    """sample_Q returns float32 like the reference; results of 2^24 values and more are converted on a few threads."""
    from qbm_b200.sampler import _as_float32
    rng = np.random.default_rng(3)
    for shape in ((7, 33), (4099, 4096)):
        a = rng.integers(0, 2, shape).astype(np.int8)
        b = _as_float32(a)
        assert b.dtype == np.float32 and b.shape == a.shape and np.array_equal(b, a.astype(np.float32))
