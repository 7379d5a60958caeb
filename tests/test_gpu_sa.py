"""GPU parity tests of the sampler path (K0-K3) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import random_qubo

pytestmark = pytest.mark.gpu


def _prep(qbm, Q, num_sweeps):
    h, J, off = qbm.ising.qubo_to_ising(Q)
    br = qbm.ising.default_beta_range(h, J)
    betas, spb = qbm.ising.beta_schedule(br, num_sweeps)
    return h[0].astype(np.float32), J[0].astype(np.float32), betas[0].astype(np.float32), spb


def test_device_philox_matches_oracle(qbm, oracle, cuda):
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2 ** 32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, size=(4096, 2), dtype=np.uint64).astype(np.uint32)
    ctr[0] = 0; key[0] = 0
    ctr[1] = 0xFFFFFFFF; key[1] = 0xFFFFFFFF
    c = torch.from_numpy(ctr.view(np.int32)).to(cuda)
    k = torch.from_numpy(key.view(np.int32)).to(cuda)
    o = torch.empty_like(c)
    L = qbm._lib.load()
    qbm._lib.check(L.qbm_test_philox(c.data_ptr(), k.data_ptr(), o.data_ptr(), 4096, None))
    torch.cuda.synchronize()
    got = o.cpu().numpy().view(np.uint32)
    # Random123 known-answer vectors
    assert got[0].tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert got[1].tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    for i in range(0, 4096, 37):
        assert got[i].tolist() == oracle.philox4x32_10(ctr[i], key[i]).tolist()


def test_device_neg_log_bit_exact(qbm, oracle, cuda):
    rng = np.random.default_rng(2)
    u = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32)
    u[:8] = [0, 1, 2, 2 ** 32 - 1, 2 ** 32 - 129, 2 ** 31, 3037000500, 12345]
    ud = torch.from_numpy(u.view(np.int32)).to(cuda)
    od = torch.empty(u.size, dtype=torch.float32, device=cuda)
    L = qbm._lib.load()
    qbm._lib.check(L.qbm_test_neg_log(ud.data_ptr(), od.data_ptr(), u.size, None))
    got = od.cpu().numpy()
    ref = np.array([oracle.neg_log_u32(int(v)) for v in u], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    fin = u > 0
    assert np.max(np.abs(got[fin] + np.log(u[fin].astype(np.float64) / 2.0 ** 32))) < 2.5e-6


@pytest.mark.parametrize("n,reads,sweeps,density", [
    (1, 5, 50, 1.0), (7, 16, 200, 1.0), (24, 33, 1000, 1.0), (32, 9, 300, 1.0), (33, 9, 300, 1.0),
    (34, 40, 1000, 1.0), (64, 8, 300, 1.0), (65, 8, 300, 0.5), (100, 8, 400, 1.0), (128, 8, 300, 1.0),
    (129, 8, 300, 1.0), (193, 12, 1000, 0.89), (256, 6, 200, 1.0), (300, 6, 200, 1.0), (522, 6, 1000, 1.0),
    (700, 4, 100, 1.0), (1000, 4, 100, 1.0), (1200, 3, 100, 1.0), (1500, 3, 100, 1.0), (1700, 3, 100, 1.0),
    (2048, 3, 1000, 1.0),
])
def test_trajectory_bit_exact_vs_replay(qbm, oracle, cuda, n, reads, sweeps, density):
    """The kernel's final states equal the sequential CPU replay of the reference's Metropolis rule fed
    the same Philox stream (north_star correctness criterion 2) -- every bit of every read."""
    Q = random_qubo(n, seed=19 + n, density=density)
    h, J, betas, spb = _prep(qbm, Q, sweeps)
    seed, off = 0x1234ABCD5678EF01 ^ n, 7 * n
    res = qbm.sa_sample(torch.from_numpy(J).to(cuda), torch.from_numpy(h).to(cuda), torch.from_numpy(betas).to(cuda),
                        spb, reads, seed, chain_offset=off, count=True)
    got = res.states.cpu().numpy()[0]
    ref, counters = oracle.replay_sample(J, h, betas, spb, seed, off, reads)
    assert np.array_equal(got, ref)
    acc = res.accepted.cpu().numpy().astype(np.uint64)
    assert int(acc[0]) == int(counters[0])
    assert int(acc[1]) == reads * n * len(betas) * spb


def test_trajectory_with_host_initial_states_and_batch(qbm, oracle, cuda):
    """batch_q problems with individual schedules + host (numpy RandomState) initial states."""
    B, n, reads, sweeps = 5, 41, 20, 300
    Qs = np.stack([random_qubo(n, seed=100 + b, scale=1.0 + b) for b in range(B)])
    h, J, _ = qbm.ising.qubo_to_ising(Qs)
    betas, spb = qbm.ising.beta_schedule(qbm.ising.default_beta_range(h, J), sweeps)
    init = np.stack([qbm.ising.initial_states_numpy(44 + b, reads, n) for b in range(B)])
    J32, h32, b32 = J.astype(np.float32), h.astype(np.float32), betas.astype(np.float32)
    res = qbm.sa_sample(torch.from_numpy(J32).to(cuda), torch.from_numpy(h32).to(cuda), torch.from_numpy(b32).to(cuda),
                        spb, reads, 77, init_states=torch.from_numpy(init).to(cuda))
    got = res.states.cpu().numpy()
    for b in range(B):
        ref, _ = oracle.replay_sample(J32[b], h32[b], b32[b], spb, 77, b * reads, reads, init01=init[b])
        assert np.array_equal(got[b], ref), f"problem {b}"


def test_sharding_invariance(qbm, cuda):
    """Reads keyed by their global index: two half-launches equal one full launch (multi-GPU sharding)."""
    n, reads = 150, 64
    Q = random_qubo(n, seed=5)
    h, J, betas, spb = _prep(qbm, Q, 200)
    Jd, hd, bd = (torch.from_numpy(a).to(cuda) for a in (J, h, betas))
    full = qbm.sa_sample(Jd, hd, bd, spb, reads, 9).states
    a = qbm.sa_sample(Jd, hd, bd, spb, reads // 2, 9, chain_offset=0).states
    b = qbm.sa_sample(Jd, hd, bd, spb, reads // 2, 9, chain_offset=reads // 2).states
    assert torch.equal(full[0], torch.cat([a[0], b[0]], dim=0))
    # the optional CTA rendezvous is a scheduling hint only: results are identical with it
    sync = qbm.sa_sample(Jd, hd, bd, spb, reads, 9, flags=1).states
    assert torch.equal(full, sync)
    Q5 = random_qubo(600, seed=6)
    h5, J5, b5, spb5 = _prep(qbm, Q5, 100)
    a5 = [torch.from_numpy(a).to(cuda) for a in (J5, h5, b5)]
    assert torch.equal(qbm.sa_sample(*a5, spb5, 19, 9).states, qbm.sa_sample(*a5, spb5, 19, 9, flags=1).states)


# ---- the chain-tile kernel (sa_tile.cu): same trajectories, rows shared by 16 chains -------------------
TILE, WARP = 16, 8      # qbm_sa_sample flag bits 4 / 3: force the chain-tile / the warp-per-chain kernel


@pytest.mark.parametrize("n,reads,sweeps,density", [
    (1, 5, 50, 1.0), (7, 16, 200, 1.0), (33, 17, 300, 1.0), (100, 40, 400, 1.0), (128, 16, 300, 1.0),
    (129, 9, 300, 1.0), (193, 35, 1000, 0.89), (300, 6, 200, 1.0), (522, 20, 1000, 1.0), (700, 4, 100, 1.0),
    (1000, 33, 100, 1.0), (1025, 3, 100, 1.0), (1200, 16, 100, 1.0), (1500, 3, 100, 1.0), (1800, 5, 60, 0.3),
    (2048, 18, 1000, 1.0),
])
def test_tile_kernel_bit_exact_vs_replay(qbm, oracle, cuda, n, reads, sweeps, density):
    """K1b: 16 chains per CTA in lock-step, coupling rows through the TMA ring -- every bit of every read
    equals the sequential CPU replay, for dense (hot) and sparse (cold) update paths alike."""
    Q = random_qubo(n, seed=19 + n, density=density)
    h, J, betas, spb = _prep(qbm, Q, sweeps)
    seed, off = 0x1234ABCD5678EF01 ^ n, 7 * n
    Jd, hd, bd = (torch.from_numpy(a).to(cuda) for a in (J, h, betas))
    res = qbm.sa_sample(Jd, hd, bd, spb, reads, seed, chain_offset=off, count=True, flags=TILE)
    got = res.states.cpu().numpy()[0]
    ref, counters = oracle.replay_sample(J, h, betas, spb, seed, off, reads)
    assert np.array_equal(got, ref)
    acc = res.accepted.cpu().numpy().astype(np.uint64)
    assert int(acc[0]) == int(counters[0])
    assert int(acc[1]) == reads * n * len(betas) * spb
    # the dense/sparse switch is a performance choice only: always-sparse and always-dense agree
    for pct in (1, 255):
        alt = qbm.sa_sample(Jd, hd, bd, spb, reads, seed, chain_offset=off, flags=TILE | (pct << 8)).states
        assert torch.equal(alt, res.states), f"dense threshold {pct}"


def test_tile_kernel_batch_init_states_and_shared_stream(qbm, oracle, cuda):
    B, n, reads, sweeps = 4, 150, 21, 300
    Qs = np.stack([random_qubo(n, seed=100 + b, scale=1.0 + b) for b in range(B)])
    h, J, _ = qbm.ising.qubo_to_ising(Qs)
    betas, spb = qbm.ising.beta_schedule(qbm.ising.default_beta_range(h, J), sweeps)
    init = np.stack([qbm.ising.initial_states_numpy(44 + b, reads, n) for b in range(B)])
    J32, h32, b32 = J.astype(np.float32), h.astype(np.float32), betas.astype(np.float32)
    Jd, hd, bd = (torch.from_numpy(a).to(cuda) for a in (J32, h32, b32))
    got = qbm.sa_sample(Jd, hd, bd, spb, reads, 77, init_states=torch.from_numpy(init).to(cuda), flags=TILE).states.cpu().numpy()
    for b in range(B):
        ref, _ = oracle.replay_sample(J32[b], h32[b], b32[b], spb, 77, b * reads, reads, init01=init[b])
        assert np.array_equal(got[b], ref), f"problem {b}"
    # flag bit 1: every problem sees the stream of read r (the reference's fixed per-call seed)
    got = qbm.sa_sample(Jd, hd, bd, spb, reads, 77, chain_offset=5, flags=TILE | 2).states.cpu().numpy()
    for b in range(B):
        ref, _ = oracle.replay_sample(J32[b], h32[b], b32[b], spb, 77, 5, reads)
        assert np.array_equal(got[b], ref), f"problem {b} (shared stream)"
    # both kernels agree with each other on the default path too
    a = qbm.sa_sample(Jd, hd, bd, spb, reads, 3, flags=TILE).states
    w = qbm.sa_sample(Jd, hd, bd, spb, reads, 3, flags=WARP).states
    assert torch.equal(a, w)


# ---- the multi-chain warp kernel (sa_multi.cu): a warp anneals T chains and shares row loads ---------------
MULTI = 32      # qbm_sa_sample flag bit 5


@pytest.mark.parametrize("n,reads,sweeps,density", [
    (129, 9, 300, 1.0), (193, 35, 1000, 0.89), (256, 8, 200, 1.0), (300, 6, 200, 1.0), (400, 5, 150, 1.0),
    (522, 21, 1000, 1.0), (700, 7, 100, 1.0), (1000, 9, 100, 1.0), (1025, 3, 100, 1.0), (1200, 5, 100, 1.0),
    (1500, 3, 100, 1.0), (1800, 5, 60, 0.3), (2048, 7, 1000, 1.0),
])
def test_multi_kernel_bit_exact_vs_replay(qbm, oracle, cuda, n, reads, sweeps, density):
    """K1c: T chains per warp with shared coupling-row loads -- every bit of every read equals the sequential
    CPU replay (read counts that do not divide by T included)."""
    Q = random_qubo(n, seed=19 + n, density=density)
    h, J, betas, spb = _prep(qbm, Q, sweeps)
    seed, off = 0x1234ABCD5678EF01 ^ n, 7 * n
    Jd, hd, bd = (torch.from_numpy(a).to(cuda) for a in (J, h, betas))
    res = qbm.sa_sample(Jd, hd, bd, spb, reads, seed, chain_offset=off, count=True, flags=MULTI)
    got = res.states.cpu().numpy()[0]
    ref, counters = oracle.replay_sample(J, h, betas, spb, seed, off, reads)
    assert np.array_equal(got, ref)
    acc = res.accepted.cpu().numpy().astype(np.uint64)
    assert int(acc[0]) == int(counters[0])
    assert int(acc[1]) == reads * n * len(betas) * spb


def test_multi_kernel_batch_init_states_and_shared_stream(qbm, oracle, cuda):
    B, n, reads, sweeps = 4, 150, 21, 300
    Qs = np.stack([random_qubo(n, seed=100 + b, scale=1.0 + b) for b in range(B)])
    h, J, _ = qbm.ising.qubo_to_ising(Qs)
    betas, spb = qbm.ising.beta_schedule(qbm.ising.default_beta_range(h, J), sweeps)
    init = np.stack([qbm.ising.initial_states_numpy(44 + b, reads, n) for b in range(B)])
    J32, h32, b32 = J.astype(np.float32), h.astype(np.float32), betas.astype(np.float32)
    Jd, hd, bd = (torch.from_numpy(a).to(cuda) for a in (J32, h32, b32))
    got = qbm.sa_sample(Jd, hd, bd, spb, reads, 77, init_states=torch.from_numpy(init).to(cuda), flags=MULTI).states.cpu().numpy()
    for b in range(B):
        ref, _ = oracle.replay_sample(J32[b], h32[b], b32[b], spb, 77, b * reads, reads, init01=init[b])
        assert np.array_equal(got[b], ref), f"problem {b}"
    got = qbm.sa_sample(Jd, hd, bd, spb, reads, 77, chain_offset=5, flags=MULTI | 2).states.cpu().numpy()
    for b in range(B):
        ref, _ = oracle.replay_sample(J32[b], h32[b], b32[b], spb, 77, 5, reads)
        assert np.array_equal(got[b], ref), f"problem {b} (shared stream)"


@pytest.mark.parametrize("flags", [0, 16, 32])
def test_sweeps_per_beta_and_single_read(qbm, oracle, cuda, flags):
    """num_sweeps > 1000 -> several sweeps per beta (neal: max(1, num_sweeps // 1000)); one read; all three kernels."""
    n = 140
    Q = random_qubo(n, seed=8)
    h, J, betas, spb = _prep(qbm, Q, 2500)
    assert spb == 2 and len(betas) == 1250
    Jd, hd, bd = (torch.from_numpy(a).to(cuda) for a in (J, h, betas))
    for reads in (1, 3):
        got = qbm.sa_sample(Jd, hd, bd, spb, reads, 5, chain_offset=11, flags=flags).states.cpu().numpy()[0]
        ref, _ = oracle.replay_sample(J, h, betas, spb, 5, 11, reads)
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("n,reads,sweeps,pct", [(2048, 20, 60, 0), (1800, 35, 40, 0), (1930, 17, 50, 97), (2048, 16, 30, 1),
                                                (2000, 5, 25, 100), (897, 19, 60, 0), (1024, 33, 40, 30), (1290, 17, 40, 0),
                                                (1536, 20, 50, 0), (1700, 5, 40, 80),
                                                (300, 19, 60, 30), (512, 35, 60, 0), (522, 20, 50, 0), (700, 10, 40, 60)])
def test_two_phase_schedule_is_the_same_trajectory(qbm, oracle, cuda, n, reads, sweeps, pct):
    """n > 256 (3..6 and 8..16 windows): the chain-tile kernel anneals the hot sweeps and hands fields / spins / sweep counters
    to the warp-per-chain kernel (the default above QBM_TWO_PHASE_MIN_N, flag bit 7 elsewhere).  Same states as with the hand-over switched off
    (flag bit 6), for any threshold (bits 16..23: 1 % = the tile kernel does everything, 100 % = it hands over after its
    first sweep), partial tiles, two problems with their own schedules, host initial states -- and equal to the replay
    oracle."""
    Qs = np.stack([random_qubo(n, seed=400 + n + b, scale=1.0 + b) for b in range(2)])
    h, J, _ = qbm.ising.qubo_to_ising(Qs)
    betas, spb = qbm.ising.beta_schedule(qbm.ising.default_beta_range(h, J), sweeps)
    J32, h32, b32 = J.astype(np.float32), h.astype(np.float32), betas.astype(np.float32)
    args = [torch.from_numpy(a).to(cuda) for a in (J32, h32, b32)]
    init = np.stack([qbm.ising.initial_states_numpy(3 + b, reads, n) for b in range(2)])
    for initd in (None, torch.from_numpy(init).to(cuda)):
        two = qbm.sa_sample(*args, spb, reads, 31, chain_offset=9, init_states=initd, count=True, flags=128 | (pct << 16))
        one = qbm.sa_sample(*args, spb, reads, 31, chain_offset=9, init_states=initd, count=True, flags=64)
        assert torch.equal(two.states, one.states)
        assert torch.equal(two.accepted, one.accepted)          # accepted flips and proposals add up over the two kernels
    got = two.states.cpu().numpy()
    ref, _ = oracle.replay_sample(J32[1], h32[1], b32[1], spb, 31, 9 + reads, 2, init01=init[1][:2])
    assert np.array_equal(got[1][:2], ref)


@pytest.mark.parametrize("n", [900, 1290, 1650, 1792])
def test_kernels_agree_where_their_row_layouts_differ(qbm, cuda, n):
    """8..14 windows: the default kernel stores rows rotated by their own window, the chain-tile and chains-per-warp kernels
    do not (and pad to 16 windows above 12) -- same states from all three, with host initial states and a batch of 2."""
    reads = 6
    Qs = np.stack([random_qubo(n, seed=300 + n + b) for b in range(2)])
    h, J, _ = qbm.ising.qubo_to_ising(Qs)
    betas, spb = qbm.ising.beta_schedule(qbm.ising.default_beta_range(h, J), 60)
    init = np.stack([qbm.ising.initial_states_numpy(7 + b, reads, n) for b in range(2)])
    args = [torch.from_numpy(a.astype(np.float32)).to(cuda) for a in (J, h, betas)]
    initd = torch.from_numpy(init).to(cuda)
    base = qbm.sa_sample(*args, spb, reads, 21, init_states=initd).states
    for flags in (16, 32, 1):           # chain-tile, chains-per-warp, default kernel with the CTA rendezvous
        assert torch.equal(base, qbm.sa_sample(*args, spb, reads, 21, init_states=initd, flags=flags).states), flags
    assert not torch.equal(base[0], base[1])


@pytest.mark.parametrize("n,R,B", [(1, 3, 1), (24, 100, 3), (34, 77, 2), (193, 130, 1), (522, 65, 1), (2048, 70, 1)])
def test_energies_vs_oracle(qbm, oracle, cuda, n, R, B):
    rng = np.random.default_rng(n)
    Qs = np.stack([random_qubo(n, seed=n + b) for b in range(B)])
    if n == 24:
        Qs[1] = rng.uniform(-1, 1, (n, n))            # a full (non-triangular) matrix is legal input
    X = (rng.random((B, R, n)) < 0.5).astype(np.int8)
    got = qbm.qubo_energies(torch.from_numpy(Qs).to(cuda), torch.from_numpy(X).to(cuda)).cpu().numpy()
    for b in range(B):
        ref = oracle.qubo_energies(Qs[b], X[b])
        assert np.allclose(got[b], ref, rtol=1e-12, atol=1e-9 * n)   # criterion: <= 1e-6 relative


@pytest.mark.parametrize("n,R,B", [(1, 1, 1), (24, 100, 4), (34, 31, 2), (193, 1000, 2), (522, 100, 1), (300, 4100, 1)])
def test_phase_stats_exact(qbm, cuda, n, R, B):
    rng = np.random.default_rng(n + R)
    X = (rng.random((B, R, n)) < rng.random((B, 1, n))).astype(np.int8)
    mean, sec = qbm.phase_stats(torch.from_numpy(X).to(cuda))
    Xf = X.astype(np.float64)
    ref_mean = Xf.mean(axis=1)
    ref_sec = np.einsum("bri,brj->bij", Xf, Xf) / R
    assert np.array_equal(mean.cpu().numpy(), ref_mean)          # bit-identical to numpy's float64 means
    assert np.array_equal(sec.cpu().numpy(), ref_sec)
    mean2, none = qbm.phase_stats(torch.from_numpy(X).to(cuda), second=False)
    assert none is None and torch.equal(mean, mean2)


def test_qubo_to_ising_device_matches_host(qbm, cuda):
    for n, B in [(1, 1), (7, 2), (24, 7), (129, 1), (193, 2), (300, 1), (522, 1), (1000, 1), (2048, 1)]:
        Qs = np.stack([random_qubo(n, seed=3 * n + b, density=0.7) for b in range(B)])
        Qs[0, 0, 0] = 0.0
        h, J, off = qbm.ising.qubo_to_ising(Qs)
        br = qbm.ising.default_beta_range(h, J)
        Jd, hd, od, rd = qbm.qubo_to_ising_device(torch.from_numpy(Qs).to(cuda))
        # bit-identical to the float64 host formulas (numpy's pairwise summation order is reproduced on the device)
        assert np.array_equal(Jd.cpu().numpy(), J.astype(np.float32))
        assert np.array_equal(hd.cpu().numpy(), h.astype(np.float32))
        assert np.allclose(od.cpu().numpy(), off, rtol=1e-13, atol=1e-12)          # information only
        r = rd.cpu().numpy()
        if n > 1:
            assert np.array_equal(qbm.ising.beta_range_from_reductions(r[:, 0], r[:, 1]), br)
    Z = torch.zeros((1, 3, 3), dtype=torch.float64, device=cuda)
    _, _, _, rz = qbm.qubo_to_ising_device(Z)
    assert rz.cpu().numpy().tolist() == [[0.0, 0.0]]


@pytest.mark.parametrize("n,reads,sweeps", [(34, 16, 300), (193, 8, 200), (1100, 3, 60)])
def test_trajectory_from_the_oracles_own_glue(qbm, oracle, cuda, n, reads, sweeps):
    """The other trajectory tests prepare h, J and the schedule with the product's `qbm.ising`; here every input of the
    kernel comes from the ORACLE's restatement of the dimod / neal glue (BINARY -> SPIN, legacy beta range, geometric
    schedule: oracle.py), so a regression of the product's host logic on the GPU box cannot hide behind itself: the product's
    whole host-buffer call (`sample_qubo_batch`: device-side conversion + schedule + kernels) must return the states the
    replay oracle computes from the oracle's own inputs."""
    Q = random_qubo(n, seed=91 + n)
    h64, J64, offset, irow, icol, jval = oracle.qubo_to_ising(Q)
    betas64, spb = oracle.beta_schedule(oracle.default_beta_range(h64, jval, irow, icol), sweeps)
    J32 = np.ascontiguousarray(J64, dtype=np.float32)
    h32 = np.ascontiguousarray(h64, dtype=np.float32)
    b32 = np.ascontiguousarray(betas64, dtype=np.float32)
    seed = 4242 + n
    ref, _ = oracle.replay_sample(J32, h32, b32, spb, seed, 0, reads)
    got, energies, info = qbm.sample_qubo_batch(Q, reads, sweeps, seed=seed, initial_states_generator="philox", device=cuda)
    assert np.array_equal(got[0], ref)
    assert np.allclose(energies[0], oracle.qubo_energies(Q, ref), rtol=1e-12, atol=1e-9)
    assert abs(float(np.ravel(info["offset"])[0]) - float(offset)) <= 1e-9 * max(1.0, abs(float(offset)))
