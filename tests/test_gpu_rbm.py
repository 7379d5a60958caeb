"""GPU: the tcgen05 TF32 GEMM and the ClassificationRBM path (boundary B3) against a plain PyTorch
fp32 reference of the same op, the numpy oracle and the golden fixture from the reference's own class.
Tolerances reflect TF32 operands (10-bit mantissa) with fp32 accumulation (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-toward-zero to 10 mantissa bits (what the tensor core reads from fp32 operands)."""
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M_,N,K", [(128, 128, 32), (256, 500, 784), (784, 500, 256), (100, 36, 64), (7, 12, 8),
                                    (300, 129, 260), (1024, 1024, 512)])
def test_gemm_tf32_vs_torch(qbm, cuda, M_, N, K):
    g = torch.Generator(device="cpu").manual_seed(M_ + N + K)
    A = torch.randn(M_, K, generator=g).to(cuda)
    B = torch.randn(N, K, generator=g).to(cuda)
    C, Ct = qbm.gemm_tf32(A, B, want_transposed=True)
    ref32 = A.double() @ B.double().T
    # exact model of the tensor core inputs: products of tf32-truncated operands, fp32 accumulation
    ref_tf = (tf32(A).double() @ tf32(B).double().T)
    err_model = (C.double() - ref_tf).abs().max().item()
    scale = ref32.abs().max().item()
    assert err_model <= 2e-5 * scale + 1e-4, f"differs from the tf32 model by {err_model}"
    assert (C.double() - ref32).abs().max().item() <= 4e-3 * np.sqrt(K)          # tf32 tolerance vs fp32 math
    assert torch.equal(Ct, C.T.contiguous())
    # fused epilogue: bias + sigmoid, and SGD accumulate
    bias = torch.randn(N, generator=g).to(cuda)
    S = qbm.gemm_tf32(A, B, alpha=0.05, bias=bias, act=1)
    assert torch.allclose(S, torch.sigmoid(0.05 * ref_tf.float() + bias), atol=2e-4)
    C0 = torch.randn(M_, N, generator=g).to(cuda)
    acc = qbm.gemm_tf32(A, B, alpha=-0.01, beta=1.0, Cin=C0)
    assert torch.allclose(acc, C0 - 0.01 * ref_tf.float(), atol=1e-4 * max(1.0, scale * 0.01))


def _model_from_golden(qbm, g):
    m = qbm.B200ClassificationRBM(64, 32, k=1, num_classes=10, learning_rate=float(g["lr"]), seed=42)
    m.weights = torch.from_numpy(g["W0"]); m.class_weights = torch.from_numpy(g["U0"])
    m.visible_bias = torch.from_numpy(g["bv0"]).to(m.device); m.hidden_bias = torch.from_numpy(g["bh0"]).to(m.device)
    m.class_bias = torch.from_numpy(g["bc0"]).to(m.device)
    return m


def test_rbm_initialisation_matches_reference(qbm, cuda):
    """Same RNG protocol as ClassificationRBM.__init__ (:14-15, :26-30)."""
    g = np.load(os.path.join(G, "rbm_discriminative.npz"))
    m = qbm.B200ClassificationRBM(64, 32, k=1, num_classes=10, learning_rate=0.05, seed=42)
    assert np.array_equal(m.weights.cpu().numpy(), g["W0"])
    assert np.array_equal(m.visible_bias.cpu().numpy(), g["bv0"])
    assert m.class_weights.shape == (10, 32) and float(m.class_weights.abs().sum()) == 0.0
    assert m.weights.shape == (64, 32) and m.hidden_bias.shape == (32,) and m.class_bias.shape == (10,)


def test_rbm_primitives_vs_golden(qbm, cuda):
    g = np.load(os.path.join(G, "rbm_discriminative.npz"))
    m = _model_from_golden(qbm, g)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    onehot = torch.nn.functional.one_hot(y[0], 10).float()
    assert np.allclose(m.sample_hidden(x[0], onehot).cpu().numpy(), g["ph0"], atol=2e-3)
    assert np.allclose(m.sample_visible(torch.from_numpy(g["hbin"])).cpu().numpy(), g["pv0"], atol=2e-3)
    assert np.allclose(m.sample_class(torch.from_numpy(g["hbin"])).cpu().numpy(), g["pc0"], atol=1e-5)
    assert np.allclose(m.sample_class_given_x(x[0]).cpu().numpy(), g["pyx0"], atol=3e-3)


def test_rbm_discriminative_steps_vs_golden(qbm, cuda):
    """R4/R5: two steps of discriminative_training against the reference's own class (torch CPU fp32)."""
    g = np.load(os.path.join(G, "rbm_discriminative.npz"))
    m = _model_from_golden(qbm, g)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    for s in range(2):
        err, pred, probs = m.discriminative_training(x[s], y[s])
        assert probs.shape == (16, 10)
        assert np.allclose(probs.cpu().numpy(), g[f"probs{s}"], atol=3e-3)
        agree = np.mean(pred.cpu().numpy() == g[f"pred{s}"])
        assert agree >= 0.9
        assert abs(float(err) - float(g[f"err{s}"])) < 2e-3
        assert np.allclose(m.weights.cpu().numpy(), g[f"W{s + 1}"], atol=2e-4)
        assert np.allclose(m.class_weights.cpu().numpy(), g[f"U{s + 1}"], atol=2e-4)
        assert np.allclose(m.hidden_bias.cpu().numpy(), g[f"bh{s + 1}"], atol=2e-4)
        assert np.allclose(m.class_bias.cpu().numpy(), g[f"bc{s + 1}"], atol=2e-4)
        assert np.array_equal(m.visible_bias.cpu().numpy(), g[f"bv{s + 1}"])
        assert torch.equal(m._Wt[:, :64], m._W[:, :32].T)               # W^T kept in sync
    with pytest.raises(ValueError):
        m.discriminative_training(x[0][:1], y[0][:1])


def synthetic_images(n, seed, V=784, C=10):
    """binarised images, pixel-on probability 0.2 + 0.6 * template_c[pixel]."""
    rng = np.random.default_rng(19)
    templates = (rng.random((C, V)) < 0.5).astype(np.float32)
    rng = np.random.default_rng(seed)
    y = rng.integers(0, C, n)
    x = (rng.random((n, V)) < 0.2 + 0.6 * templates[y]).astype(np.float32)
    return x, y


def test_rbm_training_accuracy_within_1pp_of_cpu_reference_math(qbm, cuda):
    """Config C2 shapes (784 + 10 visible, 500 hidden, batch 256): a few discriminative steps on synthetic
    images; test accuracy within 1 pp of the same steps done by the oracle in fp32 numpy."""
    V, H, C, B = 784, 500, 10, 256
    xtr, ytr = synthetic_images(1024, 1)
    xte, yte = synthetic_images(512, 2)
    m = qbm.B200ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=0.05, seed=42)
    W, U = m.weights.cpu().numpy().copy(), m.class_weights.cpu().numpy().copy()
    bv, bh, bc = m.visible_bias.cpu().numpy().copy(), m.hidden_bias.cpu().numpy().copy(), m.class_bias.cpu().numpy().copy()
    for s in range(4):
        xb, yb = xtr[s * B:(s + 1) * B], ytr[s * B:(s + 1) * B]
        m.discriminative_training(torch.from_numpy(xb), torch.from_numpy(yb))
        new, *_ = M.rbm_discriminative_step(W, U, bv, bh, bc, xb, yb, 0.05)
        W, U, bv, bh, bc = (new[k].astype(np.float32) for k in ("W", "U", "b_v", "b_h", "b_c"))
    acc_gpu = float((m.predict(torch.from_numpy(xte)).cpu().numpy() == yte).mean())
    acc_ref = float((M.rbm_class_given_x(W, U, bh, bc, xte).argmax(axis=1) == yte).mean())
    assert acc_ref > 0.5, "synthetic task should be learnable"
    assert abs(acc_gpu - acc_ref) <= 0.01
    assert np.allclose(m.weights.cpu().numpy(), W, atol=5e-4)


def test_rbm_epochs_accuracy_within_1pp_discriminative_and_cd1(qbm, cuda):
    """north_star criterion 4 at the C2 shapes (784 + 10 visible, 500 hidden, batch 256) over training RUNS, not single
    steps: 3 epochs over 2560 synthetic images (30 steps) in each training mode, then test accuracy on 2000 images.
      discriminative  (ClassificationRBM.py:101-146): the GPU run against the same steps in the float32 numpy oracle
      CD-1            (the composition of :43-60): the GPU run against the float64 replay oracle fed the kernel's
                      own Philox streams step by step
    Both arms start from the same initial draws and see the same minibatches; |accuracy difference| <= 1 pp."""
    V, H, C, B, steps = 784, 500, 10, 256, 30
    xtr, ytr = synthetic_images(2560, 1)
    xte, yte = synthetic_images(2000, 2)
    for mode, lr in (("discriminative", 0.05), ("cd1", 0.05)):
        m = qbm.B200ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=lr, seed=42)
        W, U = m.weights.cpu().numpy().astype(np.float64), m.class_weights.cpu().numpy().astype(np.float64)
        bv, bh, bc = (t.cpu().numpy().astype(np.float64) for t in (m.visible_bias, m.hidden_bias, m.class_bias))
        for s in range(steps):
            i = (s % 10) * B
            xb, yb = xtr[i:i + B], ytr[i:i + B]
            if mode == "discriminative":
                m.discriminative_training(torch.from_numpy(xb), torch.from_numpy(yb))
                new, *_ = M.rbm_discriminative_step(W.astype(np.float32), U.astype(np.float32), bv.astype(np.float32),
                                                    bh.astype(np.float32), bc.astype(np.float32), xb, yb, np.float32(lr))
            else:
                m.cd1_training(torch.from_numpy(xb), torch.from_numpy(yb))
                new, _ = M.rbm_cd1_step_replay(W, U, bv, bh, bc, xb, yb, lr, 42, s)
            W, U, bv, bh, bc = (np.asarray(new[k], dtype=np.float64) for k in ("W", "U", "b_v", "b_h", "b_c"))
        acc_gpu = float((m.predict(torch.from_numpy(xte)).cpu().numpy() == yte).mean())
        acc_ref = float((M.rbm_class_given_x(W, U, bh, bc, xte.astype(np.float64)).argmax(axis=1) == yte).mean())
        assert acc_ref > 0.5, (mode, acc_ref)                       # the synthetic task is learnable in this mode
        assert abs(acc_gpu - acc_ref) <= 0.01 + 1e-9, (mode, acc_gpu, acc_ref)


def test_rbm_cd1_step_statistics(qbm, cuda):
    """CD-1 composition: the update equals lr/B * (v0^T ph0 - v1^T ph1) for SOME valid Bernoulli draws: check
    the deterministic parts exactly (positive phase) and the sampled parts statistically."""
    V, H, C, B = 784, 500, 10, 256
    x, y = synthetic_images(B, 3)
    m = qbm.B200ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=0.1, seed=7)
    m.class_weights = torch.randn(C, H) * 0.05
    W0 = m.weights.clone(); U0 = m.class_weights.clone()
    bv0, bh0, bc0 = m.visible_bias.clone(), m.hidden_bias.clone(), m.class_bias.clone()
    xt, yt = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    onehot = torch.nn.functional.one_hot(yt, C).float()
    ph0 = torch.sigmoid(xt @ W0 + bh0 + onehot @ U0)
    m.cd1_training(xt, yt)
    dW = (m.weights - W0) * B / 0.1                       # = v0^T ph0 - v1^T ph1
    dbv = (m.visible_bias - bv0) * B / 0.1                # = sum(v0 - v1)
    dbh = (m.hidden_bias - bh0) * B / 0.1                 # = sum(ph0 - ph1)
    # expectation of the negative phase under the model (mean-field estimate): v1 ~ sigmoid(h0 W^T + b_v)
    pv1 = torch.sigmoid(ph0 @ W0.T + bv0)
    assert torch.isfinite(m.weights).all()
    # visible bias gradient: sum(v0) - sum(v1), v1 Bernoulli with mean ~ pv1 (loose 6-sigma band per pixel)
    exp_dbv = xt.sum(0) - pv1.sum(0)
    sd = torch.sqrt((pv1 * (1 - pv1)).sum(0) + 1.0)
    assert ((dbv - exp_dbv).abs() <= 6 * sd + 25).float().mean() > 0.98
    # hidden bias: ph1 is a probability, so |dbh| <= B and sign structure follows ph0 - ph1
    assert dbh.abs().max() <= B + 1e-3
    # weights: positive part must be present: dW + v1^T ph1 = v0^T ph0  =>  dW <= v0^T ph0 elementwise (+ tf32 slack)
    pos = xt.T @ ph0
    assert (dW <= pos + 0.5).all() and (dW >= pos - B - 0.5).all()
    assert torch.equal(m._Wt[:, :V], m._W[:, :H].T)
    # reproducible: same seed, same step -> same update
    m2 = qbm.B200ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=0.1, seed=7)
    m2.class_weights = U0
    m2.cd1_training(xt, yt)
    assert torch.equal(m2.weights, m.weights) and torch.equal(m2.class_bias, m.class_bias)


def test_rbm_weight_setters_do_not_alias_their_own_storage(qbm, cuda):
    """`m.weights += d` / `m.weights = m.weights` (the reference's update_weights pattern, ClassificationRBM.py:88-99) hand
    the getter's view back to the setter; W, W^T and U must come out updated, not zeroed."""
    m = qbm.B200ClassificationRBM(30, 18, 1, num_classes=3, learning_rate=0.1, seed=2)
    m.class_weights = torch.randn(3, 18)
    W0, U0 = m.weights.clone(), m.class_weights.clone()
    m.weights = m.weights
    m.class_weights = m.class_weights
    assert torch.equal(m.weights, W0) and torch.equal(m.class_weights, U0)
    d = torch.randn(30, 18, device=cuda)
    m.weights += d
    m.class_weights *= 0.5
    assert torch.equal(m.weights, W0 + d) and torch.equal(m.class_weights, U0 * 0.5)
    assert torch.equal(m._Wt[:, :30], m._W[:, :18].t())
    m.weights.add_(1.0)                                     # in-place through the view: W^T needs an explicit refresh
    m.sync_transposed_weights()
    assert torch.equal(m._Wt[:, :30], (W0 + d + 1.0).t())
    with pytest.raises(ValueError):
        m.weights = torch.zeros(18, 30)


def _cd1_intermediates(qbm, m, B):
    """Views of the CD-1 intermediates the step leaves in the model's workspace (qbm_rbm_workspace_layout)."""
    import ctypes
    V, H, C = m.num_visible, m.num_hidden, m.num_classes
    off = (ctypes.c_longlong * 12)()
    qbm._lib.check(qbm._lib.load().qbm_rbm_workspace_layout(B, V, H, C, off))
    ld4 = lambda c: (c + 3) & ~3
    ws = m._ws

    def mat(i, rows, cols):
        return ws[off[i]:off[i] + rows * ld4(cols)].view(rows, ld4(cols))[:, :cols].cpu().numpy()

    y1 = ws[off[11]:off[11] + B].view(torch.int32).cpu().numpy().astype(np.int64)
    return dict(p0=mat(4, B, H), h0=mat(6, B, H), v1=mat(7, B, V), p1t=mat(9, H, B), pc=mat(10, B, C), y1=y1)


@pytest.mark.parametrize("V,H,C,B", [(784, 500, 10, 256), (64, 48, 4, 32), (130, 70, 7, 50)])
def test_rbm_cd1_step_replays_the_kernel_draws(qbm, cuda, V, H, C, B):
    """CD-1 replay oracle (oracle/model_oracle.py::rbm_cd1_step_replay: the composition of ClassificationRBM.py:43-60 that
    SURVEY.md section 8a specifies, in float64, fed the kernel's own Philox streams (seed, step*4 + stream)):
      * ph0, p(y|h0), ph1 equal the oracle's probabilities at TF32 tolerance
      * every Bernoulli / categorical draw the kernel made is the draw its OWN probability and the Philox uniform imply,
        exactly (h0 against the exported ph0, y1 against the exported p(y|h0)); v1, whose probabilities stay inside the GEMM
        epilogue, equals the oracle's draw from float64 probabilities everywhere except where |u - p| is inside the TF32
        error band, and the whole trajectory (h0, v1, y1) equals the float64 replay except inside such bands
      * the parameter update equals R5 applied to the kernel's own phase samples at TF32 tolerance."""
    rng = np.random.default_rng(V + B)
    lr, seed = 0.1, 0x12345678
    m = qbm.B200ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=lr, seed=seed)
    m.class_weights = torch.from_numpy((rng.standard_normal((C, H)) * 0.05).astype(np.float32))
    m.hidden_bias = torch.from_numpy((rng.standard_normal(H) * 0.1).astype(np.float32)).to(cuda)
    m.class_bias = torch.from_numpy((rng.standard_normal(C) * 0.1).astype(np.float32)).to(cuda)
    v0 = (rng.random((B, V)) < 0.3).astype(np.float32)
    y0 = rng.integers(0, C, B)
    tol = 4e-3                                               # TF32 operands: |p_gpu - p_f64| band around a uniform
    for step in range(2):
        W, U = m.weights.cpu().numpy().astype(np.float64), m.class_weights.cpu().numpy().astype(np.float64)
        bv, bh, bc = (t.cpu().numpy().astype(np.float64) for t in (m.visible_bias, m.hidden_bias, m.class_bias))
        assert m._step == step
        m.cd1_training(torch.from_numpy(v0), torch.from_numpy(y0))
        torch.cuda.synchronize()
        k = _cd1_intermediates(qbm, m, B)
        new, o = M.rbm_cd1_step_replay(W, U, bv, bh, bc, v0, y0, lr, seed, step)
        # positive phase and its draw
        assert np.abs(k["p0"] - o["ph0"]).max() < tol
        assert np.array_equal(k["h0"], (o["uh"] < k["p0"]).astype(np.float32)), "h0 is not the draw of the kernel's own ph0"
        band_h = np.abs(o["uh"] - o["ph0"]) < tol
        assert np.array_equal(k["h0"][~band_h], o["h0"][~band_h].astype(np.float32)) and band_h.mean() < 0.02
        # negative phase, conditioned on the kernel's h0 (identical to the replay's unless a draw fell inside a band)
        h0 = k["h0"].astype(np.float64)
        pv1 = M.rbm_sample_visible(W, bv, h0)
        band_v = np.abs(o["uv"] - pv1) < tol
        assert np.array_equal(k["v1"][~band_v], (o["uv"] < pv1)[~band_v].astype(np.float32)) and band_v.mean() < 0.02
        pc = M.rbm_sample_class(U, bc, h0)
        assert np.abs(k["pc"] - pc).max() < tol
        assert np.array_equal(k["y1"], M.rbm_categorical_draw(k["pc"], o["uc"])), "y1 is not the draw of the kernel's own p(y|h0)"
        v1, y1 = k["v1"].astype(np.float64), k["y1"]
        ph1 = M.rbm_sample_hidden(W, U, bh, v1, np.eye(C)[y1])
        assert np.abs(k["p1t"].T - ph1).max() < tol
        if not band_h.any():                                 # no draw near a band: the whole trajectory is the float64 replay
            assert np.array_equal(h0, o["h0"])
        # parameter update from the kernel's own samples (R5)
        ref = M.rbm_cd1_update(W, U, bv, bh, bc, v0.astype(np.float64), y0, k["p0"].astype(np.float64), v1, y1, ph1, lr)
        assert np.allclose(m.weights.cpu().numpy(), ref["W"], rtol=0, atol=3e-4)
        assert np.allclose(m.class_weights.cpu().numpy(), ref["U"], rtol=0, atol=3e-4)
        assert np.allclose(m.visible_bias.cpu().numpy(), ref["b_v"], rtol=0, atol=1e-5)
        assert np.allclose(m.hidden_bias.cpu().numpy(), ref["b_h"], rtol=0, atol=3e-4)
        assert np.allclose(m.class_bias.cpu().numpy(), ref["b_c"], rtol=0, atol=1e-5)


def test_train_rbm_epoch_loop(qbm, cuda):
    """ClassificationRBM.train_rbm (src/ClassificationRBM.py:159-205): loader of (batch, labels) pairs, returns
    (loss_list, model, nll_list); the loss falls and the test accuracy beats chance on separable synthetic data."""
    rng = np.random.default_rng(0)
    templates = rng.random((4, 64)) < 0.5
    y = rng.integers(0, 4, 512)
    x = (rng.random((512, 64)) < (0.15 + 0.7 * templates[y])).astype(np.float32)
    loader = [(torch.from_numpy(x[i:i + 64]).reshape(64, 8, 8), torch.from_numpy(y[i:i + 64])) for i in range(0, 448, 64)]
    test = [(torch.from_numpy(x[448:]), torch.from_numpy(y[448:]))]
    m = qbm.B200ClassificationRBM(64, 32, 1, num_classes=4, learning_rate=0.5, seed=1)
    losses, model, nlls = m.train_rbm(loader, 12, test_loader=test)
    assert model is m and len(losses) == 12 and len(nlls) == 12
    assert nlls[-1] < nlls[0] and m.acc_per_epoch_list[-1] > 0.8
    with pytest.raises(NotImplementedError):
        m.train_rbm(loader, 1, method="generative")


@pytest.mark.parametrize("mode", ["disc", "cd1"])
def test_rbm_graph_replay_equals_eager_launches(qbm, cuda, mode):
    """From the second step of a shape on, the single-GPU steps are replayed as one CUDA graph (static buffers, CD-1 step
    counter on the device): the same kernels on the same inputs, so parameters and outputs are bit-identical to the eager
    launches -- also across a change of batch size and back, and when a bias tensor is re-assigned in between."""
    rng = np.random.default_rng(5)
    V, H, C = 96, 40, 5
    ms = [qbm.B200ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=0.05, sparse_constant=0.001, seed=11,
                                    device=cuda, use_graphs=ug) for ug in (False, True)]
    sizes = [24, 24, 24, 10, 24, 10, 10, 24]
    for step, B in enumerate(sizes):
        x = (rng.random((B, V)) < 0.3).astype(np.float32)
        y = rng.integers(0, C, B)
        outs = []
        for m in ms:
            if step == 5:
                m.hidden_bias = m.hidden_bias.clone()            # a new tensor: the captured pointers are stale
            if mode == "disc":
                loss, pred, probs = m.discriminative_training(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda))
                outs.append((loss.item(), pred.cpu().numpy(), probs.cpu().numpy()))
            else:
                m.cd1_training(x, y)
        if mode == "disc":
            assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
            assert outs[1][1].dtype == np.int64 and outs[1][2].shape == (B, C)
        for a, b in zip((ms[0].weights, ms[0].class_weights, ms[0].visible_bias, ms[0].hidden_bias, ms[0].class_bias, ms[0]._Wt),
                        (ms[1].weights, ms[1].class_weights, ms[1].visible_bias, ms[1].hidden_bias, ms[1].class_bias, ms[1]._Wt)):
            assert torch.equal(a, b), f"step {step}"
    assert any(isinstance(e, dict) for e in ms[1]._graphs.values()), "no step was captured"
    assert not ms[0]._graphs
