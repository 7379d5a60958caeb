"""CPU: pins the sampler oracle (neal restatement + replay) with known answers and exact identities."""
import numpy as np

from conftest import random_qubo


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2).tolist() == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_neg_log_u32_accuracy(oracle):
    """-ln(u/2^32): the log-domain form of the Metropolis test."""
    rng = np.random.default_rng(0)
    us = np.concatenate([rng.integers(1, 2 ** 32, 20000), 2 ** rng.integers(0, 32, 200) + rng.integers(0, 3, 200),
                         [1, 2, 3, 2 ** 32 - 1, 2 ** 32 - 300, 2 ** 31, 2 ** 31 - 1]]).astype(np.uint64)
    got = np.array([oracle.neg_log_u32(int(u)) for u in us])
    ref = -np.log(us.astype(np.float64) / 2.0 ** 32)
    # absolute error: fp32 rounding of u (2^-24 relative) plus the polynomial; tiny against the smallest dE scale
    assert np.max(np.abs(got - ref)) < 2.5e-6
    assert np.all(got >= 0.0)
    assert oracle.neg_log_u32(0) == np.inf
    assert abs(oracle.neg_log_u32(1) - 32 * np.log(2)) < 2e-6


def test_spin_energy_identity_and_offset(oracle):
    """x^T Q x == h.s + sum J s s + offset (Appendix A.2) -- ties neal's spin energies to QUBO energies."""
    rng = np.random.default_rng(3)
    for n in (1, 5, 24):
        Q = rng.uniform(-1, 1, (n, n))            # not even triangular
        h, Jsym, off, irow, icol, jv = oracle.qubo_to_ising(Q)
        X = (rng.random((50, n)) < 0.5).astype(np.int8)
        S = 2.0 * X - 1.0
        spin = S @ h + 0.5 * np.einsum("ri,ij,rj->r", S, Jsym, S) + off
        assert np.allclose(spin, oracle.qubo_energies(Q, X), rtol=1e-12, atol=1e-12)
        assert np.allclose(oracle.qubo_energies(Q, X), np.einsum("ri,ij,rj->r", X.astype(float), Q, X.astype(float)))


def test_neal_restatement_reports_qubo_energies_and_finds_ground_state(oracle):
    n = 16
    Q = random_qubo(n, seed=19)
    s, e, info = oracle.neal_sample(Q, 200, 1000, seed=19, return_info=True)
    assert s.shape == (200, n) and set(np.unique(s)) <= {0, 1}
    assert np.allclose(e, oracle.qubo_energies(Q, s), rtol=1e-12, atol=1e-12)
    X = ((np.arange(2 ** n)[:, None] >> np.arange(n)) & 1).astype(np.int8)
    gs = oracle.qubo_energies(Q, X).min()
    assert np.mean(np.abs(e - gs) < 1e-9) > 0.9
    assert info["num_betas"] == 1000 and info["sweeps_per_beta"] == 1
    c = info["counters"]
    assert int(c[0]) > 0 and int(c[0]) + int(c[1]) <= 200 * 1000 * n
    # deterministic in the seed, and the seed matters
    s2, _ = oracle.neal_sample(Q, 200, 1000, seed=19)
    s3, _ = oracle.neal_sample(Q, 200, 20, seed=20)
    assert np.array_equal(s, s2) and not np.array_equal(s, s3)


def test_legacy_beta_range_and_schedule(oracle):
    Q = np.array([[1.0, -2.0, 0.0], [0.0, 0.5, 4.0], [0.0, 0.0, -3.0]])
    h, _, _, irow, icol, jv = oracle.qubo_to_ising(Q)
    # spin biases: J01 = -0.5, J12 = 1.0 ; h = a/2 + sum b/4
    assert np.allclose(jv, [-0.5, 1.0]) and np.allclose(h, [0.0, 0.75, -0.5])
    hot, cold = oracle.default_beta_range(h, jv, irow, icol)
    assert np.isclose(hot, np.log(2) / 2.25)          # max_i |h_i| + sum |J_ij| = 0.75 + 0.5 + 1.0
    assert np.isclose(cold, np.log(100) / 0.5)        # smallest non-zero bias
    assert oracle.default_beta_range(np.zeros(3), np.zeros(0), np.zeros(0, int), np.zeros(0, int)) == [0.1, 1.0]
    b, spb = oracle.beta_schedule([hot, cold], 1000)
    assert len(b) == 1000 and spb == 1 and b[0] == hot and np.isclose(b[-1], cold)
    b, spb = oracle.beta_schedule([hot, cold], 20)
    assert len(b) == 20 and spb == 1
    b, spb = oracle.beta_schedule([hot, cold], 2500)
    assert spb == 2 and len(b) == 1250


def test_detailed_balance_at_fixed_beta(oracle):
    """At a fixed beta the restated Metropolis rule samples the Boltzmann distribution (n = 6)."""
    n = 6
    Q = random_qubo(n, seed=7)
    beta = 0.7
    s, _ = oracle.neal_sample(Q, 40000, 60, seed=5, beta_range=[beta, beta])
    X = ((np.arange(2 ** n)[:, None] >> np.arange(n)) & 1).astype(np.int8)
    # neal anneals the SPIN model with energies E_spin = E_qubo - offset: same Boltzmann weights
    E = oracle.qubo_energies(Q, X)
    p = np.exp(-beta * E); p /= p.sum()
    idx = (s.astype(int) << np.arange(n)).sum(axis=1)
    emp = np.bincount(idx, minlength=2 ** n) / len(idx)
    assert np.abs(emp - p).max() < 0.012
    # same test for the replay oracle (fp32 rule + Philox stream)
    h, Jsym, *_ = oracle.qubo_to_ising(Q)
    s2, _ = oracle.replay_sample(Jsym.astype(np.float32), h.astype(np.float32), np.full(60, beta, np.float32), 1, 11, 0, 40000)
    idx2 = (s2.astype(int) << np.arange(n)).sum(axis=1)
    emp2 = np.bincount(idx2, minlength=2 ** n) / len(idx2)
    assert np.abs(emp2 - p).max() < 0.012


def test_replay_agrees_statistically_with_neal_restatement(oracle):
    """Mean energy and ground-state hit rate of the two samplers agree within a stated tolerance."""
    n = 40
    Q = random_qubo(n, seed=11)
    s, e = oracle.neal_sample(Q, 300, 300, seed=3)
    h, Jsym, _, irow, icol, jv = oracle.qubo_to_ising(Q)
    b, spb = oracle.beta_schedule(oracle.default_beta_range(h, jv, irow, icol), 300)
    s2, _ = oracle.replay_sample(Jsym.astype(np.float32), h.astype(np.float32), b.astype(np.float32), spb, 3, 0, 300)
    e2 = oracle.qubo_energies(Q, s2)
    gs = min(e.min(), e2.min())
    assert abs(e.mean() - e2.mean()) < 0.02 * abs(gs)
    assert abs(np.mean(e < gs + 1e-9) - np.mean(e2 < gs + 1e-9)) < 0.15


def test_sample_Q_reference_linear_only_shortcut(oracle):
    Q = np.diag([1.0, -2.0, 0.0, 3.0])
    out = oracle.sample_Q_reference(Q, 5, 100, seed=44)
    assert out.shape == (5, 4) and out.dtype == np.float32
    assert out[:, 0].tolist() == [0.0] * 5 and out[:, 1].tolist() == [1.0] * 5 and out[:, 3].tolist() == [0.0] * 5
    coin = int(np.random.default_rng(44).integers(0, 2))
    assert out[:, 2].tolist() == [float(coin)] * 5
