"""Generates the golden fixtures in tests/golden/ by IMPORTING THE REFERENCE'S OWN PYTHON from
/root/reference (build container only) through oracle/ref_stubs.py.  The sampler under the
reference code is the oracle's restatement of dwave-neal (neal itself is not installable here), so
sample sets are "reference code + restated neal"; everything computed FROM a sample set (QUBO
matrices, statistics, parameter updates, RBM steps, recorded accuracies) is the reference's own
arithmetic.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import glob
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_stubs  # noqa: E402


def _samples_to_array(samples):
    return np.vstack([np.array(list(s.values())) for s in samples]).astype(np.int8)


def disc_qbm_loop(out):
    """discriminative_qbm.Disc_QBM, one-hot, C1 shapes (16 inputs, 10 labels, 24 hidden)."""
    import src.model.discriminative_qbm as D
    kw = dict(dim_input=16, num_classes=10, use_one_hot_encoding=True, n_hidden_nodes=24, restricted=False,
              sample_count=60, anneal_steps=200, beta_eff=1.0, parallelize=False, seed=19)
    np.random.seed(19)                       # Appendix B Q12: W_hh is drawn from the ambient state
    m = D.Disc_QBM(**kw)
    rng = np.random.default_rng(19)
    X = rng.random((4, 16))
    labels = np.array([3, 0, 9, 3])
    Y = np.eye(10)[labels]
    w0 = dict(W_vh=m.weights_all_visible_to_hidden.copy(), W_vo=m.weights_clamped_visible_to_output.copy(),
              W_oo=m.weights_output_output.copy(), b_h=m.biases_hidden.copy(), b_o=m.biases_output.copy(),
              W_hh=m.weights_hidden_hidden.copy())
    Qc = np.stack([m.create_qubo_matrix_from(X[i], Y[i]) for i in range(4)])
    Qu = np.stack([m.create_qubo_matrix_from(X[i]) for i in range(4)])
    Sc = np.stack([_samples_to_array(m.get_samples(X[i], label=Y[i])) for i in range(4)])
    Su = np.stack([_samples_to_array(m.get_samples(X[i])) for i in range(4)])
    stats = {}
    names = ["b_h", "b_o", "W_vh", "W_vo", "W_oo", "W_hh"]
    for i in range(4):
        from oracle.ref_stubs import _View
        rc = m.get_average_configuration([_View(r) for r in Sc[i]], X[i], [Y[i]])
        ru = m.get_average_configuration([_View(r) for r in Su[i]], X[i])
        for nm, a, b in zip(names, rc, ru):
            stats[f"stat_c_{nm}_{i}"] = np.asarray(a, dtype=np.float64)
            stats[f"stat_u_{nm}_{i}"] = np.asarray(b, dtype=np.float64)
    m.train_for_one_iteration(X, Y, 0.1, None)
    w1 = dict(W_vh=m.weights_all_visible_to_hidden, W_vo=m.weights_clamped_visible_to_output,
              W_oo=m.weights_output_output, b_h=m.biases_hidden, b_o=m.biases_output, W_hh=m.weights_hidden_hidden)
    pred = [m.predict(X[i])[0] for i in range(4)]
    np.savez_compressed(out, X=X, Y=Y, labels=labels, Qc=Qc, Qu=Qu, Sc=Sc, Su=Su, lr=0.1, seed=19,
                        sample_count=60, anneal_steps=200, pred_after=np.array(pred),
                        **{f"w0_{k}": v for k, v in w0.items()}, **{f"w1_{k}": v for k, v in w1.items()}, **stats)


def disc_qbm_faster(out):
    """faster_dqbm.Disc_QBM (what qbm_main.py runs): binary label, vectorised statistics with the
    quirks of SURVEY.md Appendix B Q1/Q2."""
    import torch
    import src.model.faster_dqbm as Fq
    np.random.seed(44)
    m = Fq.Disc_QBM(dim_input=20, num_classes=2, use_one_hot_encoding=False, n_hidden_nodes=6, restricted=False,
                    sample_count=50, anneal_steps=200, beta_eff=1.0, parallelize=False, seed=44)
    rng = np.random.default_rng(44)
    X = rng.random((5, 20))
    Y = np.array([1, 0, 1, 1, 0])
    w0 = dict(W_vh=m.weights_all_visible_to_hidden.copy(), W_vo=m.weights_clamped_visible_to_output.copy(),
              W_oo=m.weights_output_output.copy(), b_h=m.biases_hidden.copy(), b_o=m.biases_output.copy(),
              W_hh=m.weights_hidden_hidden.copy())
    Qc = np.stack([m.create_qubo_matrix_from(X[i], Y[i]) for i in range(5)])
    Qu = np.stack([m.create_qubo_matrix_from(X[i]) for i in range(5)])
    sc = m.get_samples_batch(X, label_batch=Y)
    su = m.get_samples_batch(X)
    Sc = np.stack([_samples_to_array(s) for s in sc])
    Su = np.stack([_samples_to_array(s) for s in su])
    names = ["b_h", "b_o", "W_vh", "W_vo", "W_oo", "W_hh"]
    stats = {}
    for nm, a, b in zip(names, m.get_average_configuration_batch(sc, X, Y), m.get_average_configuration_batch(su, X)):
        stats[f"stat_c_{nm}"] = np.asarray(a, dtype=np.float64)
        stats[f"stat_u_{nm}"] = np.asarray(b, dtype=np.float64)
    nll = m.compute_nll(Y, su, torch.nn.NLLLoss())
    _, loss = m.train_for_one_iteration(X, Y, 0.2, torch.nn.NLLLoss())
    w1 = dict(W_vh=m.weights_all_visible_to_hidden, W_vo=m.weights_clamped_visible_to_output,
              W_oo=m.weights_output_output, b_h=m.biases_hidden, b_o=m.biases_output, W_hh=m.weights_hidden_hidden)
    pred = [m.predict(X[i])[0] for i in range(5)]
    np.savez_compressed(out, X=X, Y=Y, Qc=Qc, Qu=Qu, Sc=Sc, Su=Su, lr=0.2, seed=44, sample_count=50, anneal_steps=200,
                        nll=nll, loss=loss, pred_after=np.array(pred),
                        **{f"w0_{k}": v for k, v in w0.items()}, **{f"w1_{k}": v for k, v in w1.items()}, **stats)


class _Recorder:
    """Wraps the model's sampler and records every (Q, samples) pair that crosses boundary B1."""

    def __init__(self, inner):
        self.inner, self.Q, self.S = inner, [], []

    def sample_Q(self, Q, num_reads):
        s = self.inner.sample_Q(Q, num_reads)
        self.Q.append(np.array(Q, dtype=np.float64)); self.S.append(np.array(s))
        return s


def convdeep(out, one_hot):
    """Conv_Deep_QBM through src/train/train.py::train_one_iteration (boundary B1)."""
    import src.model.cdqbm_state as C
    import src.train.train as T
    from src.train.pipeline import run_clamped, run_unclamped
    n_lab = 3 if one_hot else 1
    m = C.Conv_Deep_QBM(num_visible_nodes=100, num_lable_nodes=n_lab, image_shape=(10, 10), kernel_size=3,
                        pooling_size=2, pooling_type="deterministic", stride=1, sequential_layer_sizes=[12],
                        is_restricted=False, hidden_bias_type="shared", solver="SA", anneal=200, seed=44)
    rng = np.random.default_rng(77)
    X = rng.random((3, 10, 10)).astype(np.float32)
    Y = np.array([1, 0, 2]) if one_hot else np.array([1, 0, 1])
    num_reads, lr = 40, 0.05
    w0 = dict(kernel=m.kernel_weights.copy(), W_seq0=m.weights_sequential_layer[0].copy(),
              W_hy=m.weights_hidden_to_output.copy(), W_oo=m.weights_output_output.copy(),
              W_intra0=m.weights_interlayer_sequential[0].copy(), b_conv=m.biases_conv_units.copy(),
              b_seq=m.biases_sequential_units.copy(), b_out=m.biases_output.copy())
    rec = _Recorder(m.sampler)
    m.sampler = rec
    stats = {}
    names = ["b_conv", "b_seq", "b_out", "kernel", "W_intra", "W_seq", "W_hy", "W_oo"]
    probs = []
    for i in range(3):
        lab = np.eye(n_lab)[Y[i]] if one_hot else np.array([float(Y[i])])
        oc = run_clamped(m, X[i], lab, num_reads, 1.0)
        ou = run_unclamped(m, X[i], num_reads, 1.0, one_hot)
        probs.append(ou.probs)
        for tag, o, yy in (("c", oc, lab), ("u", ou, None)):
            r = T.get_average_configuration_single(m, o, X[i], y=yy)
            for nm, a in zip(names, r):
                a = a[0] if isinstance(a, list) else a
                stats[f"stat_{tag}_{nm}_{i}"] = np.asarray(a, dtype=np.float64)
    Qc = np.stack(rec.Q[0::2]); Qu = np.stack(rec.Q[1::2])
    Sc = np.stack(rec.S[0::2]).astype(np.int8); Su = np.stack(rec.S[1::2]).astype(np.int8)
    loss = T.train_one_iteration(m, X, Y, num_reads, 1.0, lr, one_hot=one_hot)
    w1 = dict(kernel=m.kernel_weights, W_seq0=m.weights_sequential_layer[0], W_hy=m.weights_hidden_to_output,
              W_oo=m.weights_output_output, W_intra0=m.weights_interlayer_sequential[0], b_conv=m.biases_conv_units,
              b_seq=m.biases_sequential_units, b_out=m.biases_output)
    np.savez_compressed(out, X=X, Y=Y, Qc=Qc, Qu=Qu, Sc=Sc, Su=Su, lr=lr, num_reads=num_reads, anneal=200, seed=44,
                        loss=loss, probs=np.stack(probs), one_hot=one_hot,
                        **{f"w0_{k}": v for k, v in w0.items()}, **{f"w1_{k}": v for k, v in w1.items()}, **stats)


def rbm(out):
    """ClassificationRBM.discriminative_training (src/ClassificationRBM.py:101-146), two steps."""
    import torch
    import src.ClassificationRBM as R
    V, H, C, B = 64, 32, 10, 16
    m = R.ClassificationRBM(V, H, k=1, num_classes=C, learning_rate=0.05, seed=42)
    g = torch.Generator().manual_seed(7)
    m.class_weights = torch.randn(C, H, generator=g) * 0.1      # the reference initialises U to zeros;
    m.hidden_bias = torch.randn(H, generator=g) * 0.1           # non-trivial values exercise every term
    m.class_bias = torch.randn(C, generator=g) * 0.1
    x = (torch.rand(2, B, V, generator=g) < 0.3).float()
    y = torch.randint(0, C, (2, B), generator=g)
    d = dict(x=x.numpy(), y=y.numpy(), lr=0.05,
             W0=m.weights.numpy().copy(), U0=m.class_weights.numpy().copy(), bv0=m.visible_bias.numpy().copy(),
             bh0=m.hidden_bias.numpy().copy(), bc0=m.class_bias.numpy().copy())
    d["ph0"] = m.sample_hidden(x[0], torch.nn.functional.one_hot(y[0], C).float()).numpy()
    hbin = (torch.rand(B, H, generator=g) < 0.5).float()
    d["hbin"] = hbin.numpy()
    d["pv0"] = m.sample_visible(hbin).numpy()
    d["pc0"] = m.sample_class(hbin).numpy()
    d["pyx0"] = m.sample_class_given_x(x[0]).numpy()
    for s in range(2):
        err, pred, probs = m.discriminative_training(x[s], y[s])
        d[f"err{s}"] = float(err); d[f"pred{s}"] = pred.numpy(); d[f"probs{s}"] = probs.numpy()
        d[f"W{s + 1}"] = m.weights.numpy().copy(); d[f"U{s + 1}"] = m.class_weights.numpy().copy()
        d[f"bh{s + 1}"] = m.hidden_bias.numpy().copy(); d[f"bc{s + 1}"] = m.class_bias.numpy().copy()
        d[f"bv{s + 1}"] = m.visible_bias.numpy().copy()
    np.savez_compressed(out, **d)


def recorded_accuracy(out):
    """Known-answer fixtures: trained weights + the accuracy the reference recorded for them
    (out/paper_data/Pneumonia_param_doku/10_hnodes).  Stored compactly: the
    image-independent off-diagonal part of the unclamped QUBO and the per-image diagonal, checked
    here against the reference's own create_qubo_matrix_from."""
    import src.model.faster_dqbm as Fq
    from src import data_loader
    (trX, trY), (vaX, vaY), (teX, teY) = data_loader.get_medmnist("src/data/medmnist/pneumoniamnist.npz")
    _, teXf, _ = data_loader.preprocess_images(trX[:2], teX)
    base = "out/paper_data/Pneumonia_param_doku/10_hnodes"
    d = {"labels": np.asarray(teY).astype(np.int8)}
    for k, run in enumerate(sorted(glob.glob(os.path.join(base, "_se*")))[:3]):
        seed = int(os.path.basename(run).split("_se")[1].split("_")[0])
        wfile = glob.glob(os.path.join(run, "e20__*.pkl"))[0]
        with open(wfile, "rb") as f:
            W = pickle.load(f)
        with open(os.path.join(run, "test_val", f"e20_h10_{seed}_testacc_auc.pkl"), "rb") as f:
            acc, auc = pickle.load(f)
        m = Fq.Disc_QBM(dim_input=784, num_classes=2, use_one_hot_encoding=False, n_hidden_nodes=10, restricted=False,
                        sample_count=100, anneal_steps=1000, beta_eff=1.0, parallelize=False, seed=seed)
        (m.weights_all_visible_to_hidden, m.weights_clamped_visible_to_output, m.biases_hidden, m.biases_output,
         m.weights_output_output, m.weights_hidden_hidden) = W
        Q0 = m.create_qubo_matrix_from(teXf[0])
        off = Q0 - np.diag(np.diag(Q0))
        diag = np.empty((len(teXf), 11))
        for i in range(len(teXf)):
            Q = m.create_qubo_matrix_from(teXf[i])
            assert np.array_equal(Q - np.diag(np.diag(Q)), off)
            diag[i] = np.diag(Q)
        d[f"off_{k}"] = off; d[f"diag_{k}"] = diag; d[f"acc_{k}"] = acc; d[f"auc_{k}"] = auc; d[f"seed_{k}"] = seed
    np.savez_compressed(out, **d)


def recorded_accuracy_all(out):
    """All 70 PneumoniaMNIST last-epoch runs with h in {4,5,6,7,8,10,12} (10 seeds each; SURVEY.md section 4 found the
    recorded (accuracy, AUC) of every one of them reproduced by the ground state of the unclamped QUBO).  Per run: the
    image-independent off-diagonal part of the unclamped QUBO and the per-image diagonal from the reference's own
    create_qubo_matrix_from, plus the recorded pair.  Diagonals are stored as float32: enough to leave every ground
    state unchanged (checked here against the float64 QUBO by exact enumeration) at half the fixture size."""
    import re
    import src.model.faster_dqbm as Fq
    from src import data_loader
    (trX, trY), (vaX, vaY), (teX, teY) = data_loader.get_medmnist("src/data/medmnist/pneumoniamnist.npz")
    _, teXf, _ = data_loader.preprocess_images(trX[:2], teX)
    teXf = np.asarray(teXf, dtype=np.float64)
    labels = np.asarray(teY).astype(np.int8).reshape(-1)
    d = {"labels": labels}
    k = 0
    for h in (4, 5, 6, 7, 8, 10, 12):
        base = f"out/paper_data/Pneumonia_param_doku/{h}_hnodes"
        for run in sorted(glob.glob(os.path.join(base, "_se*"))):
            seed = int(os.path.basename(run).split("_se")[1].split("_")[0])
            wfiles = glob.glob(os.path.join(run, "e*__*.pkl"))
            last = max(int(re.match(r"e(\d+)__", os.path.basename(f)).group(1)) for f in wfiles)
            wfile = [f for f in wfiles if os.path.basename(f).startswith(f"e{last}__")][0]
            with open(wfile, "rb") as f:
                W = pickle.load(f)
            with open(os.path.join(run, "test_val", f"e{last}_h{h}_{seed}_testacc_auc.pkl"), "rb") as f:
                acc, auc = pickle.load(f)
            restricted = len(W) == 5
            m = Fq.Disc_QBM(dim_input=784, num_classes=2, use_one_hot_encoding=False, n_hidden_nodes=h,
                            restricted=restricted, sample_count=100, anneal_steps=1000, beta_eff=1.0, parallelize=False,
                            seed=seed)
            if restricted:
                (m.weights_all_visible_to_hidden, m.weights_clamped_visible_to_output, m.biases_hidden, m.biases_output,
                 m.weights_output_output) = W
            else:
                (m.weights_all_visible_to_hidden, m.weights_clamped_visible_to_output, m.biases_hidden, m.biases_output,
                 m.weights_output_output, m.weights_hidden_hidden) = W
            n = h + 1
            Q0 = m.create_qubo_matrix_from(teXf[0])
            off = Q0 - np.diag(np.diag(Q0))
            # the diagonal of every image in one product, checked against the reference's builder on a few images
            lin = np.concatenate([np.ravel(m.biases_output), np.ravel(m.biases_hidden)])[None, :] + teXf @ np.concatenate(
                [m.weights_clamped_visible_to_output, m.weights_all_visible_to_hidden[1:]], axis=1)
            lin = lin / m.beta_eff
            for i in (0, 1, 311, 623):
                Q = m.create_qubo_matrix_from(teXf[i])
                assert np.array_equal(Q - np.diag(np.diag(Q)), off)
                assert np.allclose(np.diag(Q), lin[i], rtol=1e-12, atol=1e-12), (h, seed, i)
            diag32 = lin.astype(np.float32)
            X = ((np.arange(2 ** n)[:, None] >> np.arange(n)) & 1).astype(np.float64)
            quad = np.einsum("ri,ij,rj->r", X, off, X)
            gs64 = np.argmin(lin @ X.T + quad[None], axis=1)
            gs32 = np.argmin(diag32.astype(np.float64) @ X.T + quad[None], axis=1)
            assert np.array_equal(gs64, gs32), (h, seed)
            pred = (gs64 & 1).astype(int)
            assert abs(np.mean(pred == labels) - acc) < 1e-12, (h, seed, np.mean(pred == labels), acc)
            d[f"off_{k}"] = off; d[f"diag_{k}"] = diag32; d[f"acc_{k}"] = acc; d[f"auc_{k}"] = auc
            d[f"seed_{k}"] = seed; d[f"h_{k}"] = h; d[f"epoch_{k}"] = last
            d[f"sc_{k}"] = int(re.search(r"_sc(\d+)_", os.path.basename(run)).group(1))     # the run's sample count
            k += 1
    d["num_runs"] = k
    np.savez_compressed(out, **d)


def disc_qbm_c5(out):
    """discriminative_qbm.Disc_QBM at the C5 shapes (128 inputs, 10 one-hot labels, 512 hidden: n = 512 / 522), two images,
    100 reads x 1000 sweeps.  Arrays too large to commit (QUBOs, W_vh, W_hh, their statistics) are stored as digests
    (oracle.model_oracle.array_digest); sample sets and inputs in full."""
    import src.model.discriminative_qbm as D
    from oracle.model_oracle import array_digest as dg
    from oracle.ref_stubs import _View
    np.random.seed(77)                       # Appendix B Q12
    m = D.Disc_QBM(dim_input=128, num_classes=10, use_one_hot_encoding=True, n_hidden_nodes=512, restricted=False,
                   sample_count=100, anneal_steps=1000, beta_eff=1.0, parallelize=False, seed=77)
    rng = np.random.default_rng(77)
    X = rng.random((2, 128))
    labels = np.array([7, 2])
    Y = np.eye(10)[labels]
    names = ["b_h", "b_o", "W_vh", "W_vo", "W_oo", "W_hh"]
    cur = lambda: dict(W_vh=m.weights_all_visible_to_hidden, W_vo=m.weights_clamped_visible_to_output,
                       W_oo=m.weights_output_output, b_h=m.biases_hidden, b_o=m.biases_output, W_hh=m.weights_hidden_hidden)
    d = dict(X=X, Y=Y, labels=labels, lr=0.1, seed=77, sample_count=100, anneal_steps=1000)
    for k, v in cur().items():
        d[f"w0_{k}_dg"] = dg(v)
    d["w0_b_o"] = m.biases_output.copy(); d["w0_W_oo"] = m.weights_output_output.copy()
    Sc, Su = [], []
    for i in range(2):
        d[f"Qc_dg_{i}"] = dg(m.create_qubo_matrix_from(X[i], Y[i]))
        d[f"Qu_dg_{i}"] = dg(m.create_qubo_matrix_from(X[i]))
        Sc.append(_samples_to_array(m.get_samples(X[i], label=Y[i])))
        Su.append(_samples_to_array(m.get_samples(X[i])))
        rc = m.get_average_configuration([_View(r) for r in Sc[i]], X[i], [Y[i]])
        ru = m.get_average_configuration([_View(r) for r in Su[i]], X[i])
        for nm, a, b in zip(names, rc, ru):
            d[f"stat_c_{nm}_dg_{i}"] = dg(a); d[f"stat_u_{nm}_dg_{i}"] = dg(b)
    d["Sc"] = np.stack(Sc); d["Su"] = np.stack(Su)
    m.train_for_one_iteration(X, Y, 0.1, None)
    for k, v in cur().items():
        d[f"w1_{k}_dg"] = dg(v)
    d["w1_b_o"] = m.biases_output.copy(); d["w1_W_oo"] = m.weights_output_output.copy()
    np.savez_compressed(out, **d)


def convdeep_c3(out):
    """Conv_Deep_QBM at the C3 shapes (18x18 image, 3x3 kernel, pool 2 -> 64 pooled units, 128 sequential units, binary
    label: n = 192 / 193) through src/train/train.py, two images, 1000 reads x 1000 sweeps; large arrays as digests."""
    import src.model.cdqbm_state as C
    import src.train.train as T
    from src.train.pipeline import run_clamped, run_unclamped
    from oracle.model_oracle import array_digest as dg
    m = C.Conv_Deep_QBM(num_visible_nodes=324, num_lable_nodes=1, image_shape=(18, 18), kernel_size=3, pooling_size=2,
                        pooling_type="deterministic", stride=1, sequential_layer_sizes=[128], is_restricted=False,
                        hidden_bias_type="shared", solver="SA", anneal=1000, seed=44)
    rng = np.random.default_rng(19)
    X = rng.random((2, 18, 18)).astype(np.float32)
    Y = np.array([1, 0])
    num_reads, lr = 1000, 0.01
    cur = lambda: dict(kernel=m.kernel_weights, W_seq0=m.weights_sequential_layer[0], W_hy=m.weights_hidden_to_output,
                       W_oo=m.weights_output_output, W_intra0=m.weights_interlayer_sequential[0], b_conv=m.biases_conv_units,
                       b_seq=m.biases_sequential_units, b_out=m.biases_output)
    d = dict(X=X, Y=Y, lr=lr, num_reads=num_reads, anneal=1000, seed=44)
    for k, v in cur().items():
        d[f"w0_{k}_dg"] = dg(v)
    d["w0_kernel"] = m.kernel_weights.copy()
    rec = _Recorder(m.sampler)
    m.sampler = rec
    names = ["b_conv", "b_seq", "b_out", "kernel", "W_intra", "W_seq", "W_hy", "W_oo"]
    probs = []
    for i in range(2):
        lab = np.array([float(Y[i])])
        oc = run_clamped(m, X[i], lab, num_reads, 1.0)
        ou = run_unclamped(m, X[i], num_reads, 1.0, False)
        probs.append(ou.probs)
        for tag, o, yy in (("c", oc, lab), ("u", ou, None)):
            r = T.get_average_configuration_single(m, o, X[i], y=yy)
            for nm, a in zip(names, r):
                a = a[0] if isinstance(a, list) else a
                d[f"stat_{tag}_{nm}_dg_{i}"] = dg(a)
    for i in range(2):
        d[f"Qc_dg_{i}"] = dg(rec.Q[2 * i]); d[f"Qu_dg_{i}"] = dg(rec.Q[2 * i + 1])
    d["Sc"] = np.stack(rec.S[0:4:2]).astype(np.int8); d["Su"] = np.stack(rec.S[1:4:2]).astype(np.int8)
    d["probs"] = np.stack(probs)
    d["loss"] = T.train_one_iteration(m, X, Y, num_reads, 1.0, lr, one_hot=False)
    for k, v in cur().items():
        d[f"w1_{k}_dg"] = dg(v)
    d["w1_kernel"] = m.kernel_weights.copy()
    np.savez_compressed(out, **d)


def main():
    with ref_stubs.reference_imports():
        disc_qbm_loop(os.path.join(HERE, "disc_qbm_loop_onehot.npz"))
        disc_qbm_faster(os.path.join(HERE, "disc_qbm_faster_binary.npz"))
        convdeep(os.path.join(HERE, "convdeep_binary.npz"), one_hot=False)
        convdeep(os.path.join(HERE, "convdeep_onehot.npz"), one_hot=True)
        rbm(os.path.join(HERE, "rbm_discriminative.npz"))
        recorded_accuracy(os.path.join(HERE, "pneumonia_h10_recorded_accuracy.npz"))
        recorded_accuracy_all(os.path.join(HERE, "pneumonia_last_epoch_recorded_accuracy.npz"))
        disc_qbm_c5(os.path.join(HERE, "disc_qbm_loop_c5.npz"))
        convdeep_c3(os.path.join(HERE, "convdeep_c3.npz"))
    for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        print(f"{os.path.basename(f):45s} {os.path.getsize(f) / 1024:8.1f} KB")


if __name__ == "__main__":
    main()
