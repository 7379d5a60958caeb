"""GPU: the Conv-Deep path (K6 context kernel, batched QUBO builders, batched training step) against the
numpy oracle and the golden fixtures generated from the reference's own Python (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("shape,k,stride,pool", [((10, 10), 3, 1, 2), ((18, 18), 3, 1, 2), ((28, 28), 3, 1, 4),
                                                 ((12, 13), 5, 2, 0), ((9, 9), 2, 1, 3), ((28, 28), 3, 2, 1),
                                                 ((16, 16), 11, 1, 2)])
def test_context_kernel_bit_exact(qbm, cuda, shape, k, stride, pool):
    """conv2d_valid_stride + pooled_indices_for_input + patch gather (geometry.py:37-53, layers.py:65-84,
    train.py:188-191): feature maps are bit-identical to numpy's, so the argmin cannot differ."""
    rng = np.random.default_rng(5)
    m = qbm.ConvDeepQBM(shape[0] * shape[1], 1, image_shape=shape, kernel_size=k, pooling_size=pool, stride=stride,
                        sequential_layer_sizes=[4], hidden_bias_type="shared", seed=3)
    X = rng.random((6,) + shape).astype(np.float32)
    X[5] = np.round(X[5] * 2) / 2                    # ties inside pooling windows: first minimum wins
    fmap, pooled, patches = (t.cpu().numpy() for t in m.prepare_context_batch(X))
    kern = m.kernel_weights
    for b in range(6):
        f, pi, pt = M.convdeep_context(X[b], kern, stride, pool)
        assert np.array_equal(fmap[b], f)
        assert np.array_equal(pooled[b], pi)
        assert np.array_equal(patches[b], pt)
    assert pooled.shape[1] == m.num_pooled_units == len(pi)


def _golden_params(g, prefix):
    return dict(kernel=g[f"{prefix}_kernel"], W_seq=[g[f"{prefix}_W_seq0"]], W_intra=[g[f"{prefix}_W_intra0"]],
                W_hy=g[f"{prefix}_W_hy"], W_oo=g[f"{prefix}_W_oo"], b_conv=g[f"{prefix}_b_conv"], b_seq=g[f"{prefix}_b_seq"],
                b_out=g[f"{prefix}_b_out"])


@pytest.mark.parametrize("name,one_hot", [("convdeep_binary.npz", False), ("convdeep_onehot.npz", True)])
def test_convdeep_training_step(qbm, cuda, name, one_hot):
    """P1-P3 (pipeline.py:13-36, train.py:12-253): initial draws, both QUBO builders and one whole training
    step against the reference's arithmetic."""
    g = np.load(os.path.join(G, name))
    n_lab = 3 if one_hot else 1
    m = qbm.ConvDeepQBM(num_visible_nodes=100, num_lable_nodes=n_lab, image_shape=(10, 10), kernel_size=3, pooling_size=2,
                        pooling_type="deterministic", stride=1, sequential_layer_sizes=[12], is_restricted=False,
                        hidden_bias_type="shared", solver="SA", anneal=int(g["anneal"]), seed=int(g["seed"]))
    p0 = _golden_params(g, "w0")
    got = m.get_params()
    for k, v in p0.items():
        a, b = (got[k][0], v[0]) if isinstance(v, list) else (got[k], v)
        assert np.array_equal(a, b), f"initial {k} differs from the reference's draw"
    X, Y = g["X"], g["Y"]
    fmap, pooled, patches = m.prepare_context_batch(X)
    lab = torch.from_numpy(np.eye(n_lab)[Y] if one_hot else Y.astype(np.float64)[:, None]).to(cuda)
    assert np.allclose(m.build_qubos(fmap, pooled, lab).cpu().numpy(), g["Qc"], rtol=0, atol=1e-13)
    assert np.allclose(m.build_qubos(fmap, pooled, None).cpu().numpy(), g["Qu"], rtol=0, atol=1e-13)
    m.keep_samples = True
    R, lr = int(g["num_reads"]), float(g["lr"])
    loss = m.train_one_iteration(X, Y, R, 1.0, lr, one_hot=one_hot)
    Sc, Su = (t.cpu().numpy().astype(np.float32) for t in m.last_samples)
    assert Sc.shape == g["Sc"].shape and Su.shape == g["Su"].shape
    ref, ref_loss = M.convdeep_train_step(p0, X, Y, Sc, Su, lr, 1, 2, one_hot)     # reference arithmetic, OUR samples
    assert abs(loss - ref_loss) < 1e-6
    got = m.get_params()
    for k, v in ref.items():
        a, b = (got[k][0], v[0]) if isinstance(v, list) else (got[k], v)
        assert np.allclose(a, b, rtol=1e-6, atol=1e-7), k          # float32 sample means in the reference
    # the sample sets are statistically the reference's (same QUBOs, same schedule, same initial states)
    assert np.abs(Sc.mean(axis=1) - g["Sc"].astype(float).mean(axis=1)).mean() < 0.1
    # reference attribute names and checkpoint list (cdqbm_state.py:41-48)
    assert m.kernel_weights.shape == (3, 3) and m.weights_sequential_layer[0].shape == (16, 12)
    assert len(m.weight_objects) == 8
    probs = m.predict_proba_batch(X, R, 1.0, one_hot)
    assert probs.shape == (3, 3 if one_hot else 2) and np.allclose(probs.sum(axis=1), 1.0, atol=1e-6)


def test_convdeep_training_step_at_the_c3_shapes(qbm, cuda):
    """C3 (BASELINE config 3: 18x18 image, 3x3 kernel, pool 2 -> 64 pooled units, 128 sequential units, 1000 reads x 1000
    sweeps; n = 192 / 193) against the golden generated from Conv_Deep_QBM through src/train: initial draws and both QUBO
    builders equal the reference's (digests), the reference's own sample sets through K3 + K11 + K9 give its parameters and
    loss after one step, and a whole step on the GPU's samples equals the reference arithmetic on those samples."""
    g = np.load(os.path.join(G, "convdeep_c3.npz"))
    m = qbm.ConvDeepQBM(num_visible_nodes=324, num_lable_nodes=1, image_shape=(18, 18), kernel_size=3, pooling_size=2,
                        pooling_type="deterministic", stride=1, sequential_layer_sizes=[128], is_restricted=False,
                        hidden_bias_type="shared", solver="SA", anneal=int(g["anneal"]), seed=int(g["seed"]))
    flat = lambda p: dict(kernel=p["kernel"], W_seq0=p["W_seq"][0], W_hy=p["W_hy"], W_oo=p["W_oo"], W_intra0=p["W_intra"][0],
                          b_conv=p["b_conv"], b_seq=p["b_seq"], b_out=p["b_out"])
    p0 = m.get_params()
    for k, v in flat(p0).items():
        assert np.allclose(M.array_digest(v), g[f"w0_{k}_dg"], rtol=1e-10, atol=1e-9), f"initial {k}"
    X, Y = g["X"], g["Y"]
    fmap, pooled, patches = m.prepare_context_batch(X)
    assert m.num_pooled_units == 64 and m.n_hidden == 192
    lab = torch.from_numpy(Y.astype(np.float64)[:, None]).to(cuda)
    Qc, Qu = m.build_qubos(fmap, pooled, lab).cpu().numpy(), m.build_qubos(fmap, pooled, None).cpu().numpy()
    for i in range(2):
        assert np.allclose(M.array_digest(Qc[i]), g[f"Qc_dg_{i}"], rtol=1e-10, atol=1e-9)
        assert np.allclose(M.array_digest(Qu[i]), g[f"Qu_dg_{i}"], rtol=1e-10, atol=1e-9)
    R, lr = int(g["num_reads"]), float(g["lr"])
    loss = m.train_step_from_samples(X, Y, torch.from_numpy(g["Sc"]).to(cuda), torch.from_numpy(g["Su"]).to(cuda), lr)
    assert abs(loss - float(g["loss"])) < 1e-6
    for k, v in flat(m.get_params()).items():
        assert np.allclose(M.array_digest(v), g[f"w1_{k}_dg"], rtol=1e-6, atol=1e-6), f"after one step: {k}"
    p1 = m.get_params()
    m.keep_samples = True
    loss = m.train_one_iteration(X, Y, R, 1.0, lr)
    Sc, Su = (t.cpu().numpy().astype(np.float32) for t in m.last_samples)
    assert Sc.shape == (2, 1000, 192) and Su.shape == (2, 1000, 193)
    ref, ref_loss = M.convdeep_train_step(p1, X, Y, Sc, Su, lr, 1, 2, False)
    assert abs(loss - ref_loss) < 1e-6
    got = m.get_params()
    for k, v in ref.items():
        a, b = (got[k][0], v[0]) if isinstance(v, list) else (got[k], v)
        assert np.allclose(a, b, rtol=1e-6, atol=1e-7), k
    probs = m.predict_proba_batch(X, R, 1.0, False)
    assert probs.shape == (2, 2) and np.allclose(probs.sum(axis=1), 1.0, atol=1e-6)


def test_convdeep_restricted_no_bias_and_float64_stats(qbm, cuda):
    """is_restricted=True (no within-layer couplings), hidden_bias_type='none', two sequential layers."""
    rng = np.random.default_rng(11)
    m = qbm.ConvDeepQBM(144, 1, image_shape=(12, 12), kernel_size=3, pooling_size=2, sequential_layer_sizes=[10, 6],
                        is_restricted=True, hidden_bias_type="none", anneal=100, seed=44, stats_dtype="float64")
    X = rng.random((4, 12, 12)).astype(np.float32)
    Y = np.array([0, 1, 1, 0])
    p0 = m.get_params()
    assert p0["W_intra"] is None and m.n_hidden == 25 + 16
    fmap, pooled, _ = m.prepare_context_batch(X)
    ref_p = dict(p0, b_conv=np.zeros(1))
    for b in range(4):
        f, pi, _ = M.convdeep_context(X[b], p0["kernel"], 1, 2)
        Qu = M.convdeep_qubo(ref_p, f, pi, None, beta_eff=2.0, restricted=True)
        assert np.allclose(m.build_qubos(fmap[b:b + 1], pooled[b:b + 1], None, 2.0)[0].cpu().numpy(), Qu, rtol=0, atol=1e-13)
    m.keep_samples = True
    loss = m.train_one_iteration(X, Y, 30, 2.0, 0.1)
    Sc, Su = (t.cpu().numpy().astype(np.float32) for t in m.last_samples)
    ref, ref_loss = M.convdeep_train_step(ref_p, X, Y, Sc, Su, 0.1, 1, 2, False, restricted=True)
    assert abs(loss - ref_loss) < 1e-6
    got = m.get_params()
    for k in ("kernel", "W_hy", "W_oo", "b_seq", "b_out"):
        assert np.allclose(got[k], ref[k], rtol=1e-6, atol=1e-7), k
    for li in range(2):
        assert np.allclose(got["W_seq"][li], ref["W_seq"][li], rtol=1e-6, atol=1e-7)
    assert np.array_equal(got["b_conv"], p0["b_conv"])              # 'none': never touched
    with pytest.raises(ValueError):
        qbm.ConvDeepQBM(144, 1, image_shape=(12, 12), pooling_type="probabilistic")


@pytest.mark.parametrize("name,one_hot", [("convdeep_binary.npz", False), ("convdeep_onehot.npz", True), ("convdeep_c3.npz", False)])
def test_run_unclamped_probabilities_from_the_reference_samples(qbm, cuda, name, one_hot):
    """P1 (src/train/pipeline.py:24-36, RunOutputs.probs): the class probabilities the product derives from a sample set are
    the reference's, checked directly on the reference's OWN unclamped sample sets (recorded in the golden fixture with the
    probabilities run_unclamped returned for them) -- not only through the loss."""
    g = np.load(os.path.join(G, name))
    if "Su" not in g.files:
        pytest.skip("fixture keeps digests of the sample sets only")
    Su = g["Su"]
    one_hot = bool(g["one_hot"]) if "one_hot" in g.files else one_hot
    n_out = g["probs"].shape[1] if one_hot else 1
    m = qbm.ConvDeepQBM.__new__(qbm.ConvDeepQBM)          # only the hidden / output split of the variable layout is needed
    m.n_hidden = Su.shape[2] - n_out
    mu, _ = qbm.phase_stats(torch.from_numpy(np.ascontiguousarray(Su)).to(cuda), second=False)
    probs = m._probs(mu, one_hot).cpu().numpy()
    assert probs.shape == g["probs"].shape
    assert np.allclose(probs, g["probs"], rtol=0, atol=1e-6), np.abs(probs - g["probs"]).max()
