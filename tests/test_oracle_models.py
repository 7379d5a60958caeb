"""CPU: the numpy oracle of the model-side hot path against the golden fixtures produced by the
reference's own Python (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import model_oracle as M

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def params(g, prefix):
    return {k: g[f"{prefix}_{k}"] for k in ("W_vh", "W_vo", "W_oo", "b_h", "b_o", "W_hh")}


def test_disc_qbm_loop_onehot_matches_reference():
    g = load("disc_qbm_loop_onehot.npz")
    p0 = params(g, "w0")
    X, Y = g["X"], g["Y"]
    for i in range(4):
        assert np.array_equal(M.disc_qubo(p0, X[i], Y[i]), g["Qc"][i])
        assert np.array_equal(M.disc_qubo(p0, X[i]), g["Qu"][i])
        c = M.disc_stats_loop(g["Sc"][i], X[i], Y[i], 10, 16, 24)
        u = M.disc_stats_loop(g["Su"][i], X[i], None, 10, 16, 24)
        for nm, a, b in zip(["b_h", "b_o", "W_vh", "W_vo", "W_oo", "W_hh"], c, u):
            assert np.allclose(a, g[f"stat_c_{nm}_{i}"], rtol=0, atol=1e-15), nm
            assert np.allclose(b, g[f"stat_u_{nm}_{i}"], rtol=0, atol=1e-15), nm
    p1 = M.disc_train_step(p0, X, Y, g["Sc"], g["Su"], float(g["lr"]), "loop")
    for k, v in p1.items():
        assert np.allclose(v, g[f"w1_{k}"], rtol=0, atol=1e-14), k


def test_disc_qbm_faster_binary_matches_reference():
    g = load("disc_qbm_faster_binary.npz")
    p0 = params(g, "w0")
    X, Y = g["X"], g["Y"]
    for i in range(5):
        assert np.array_equal(M.disc_qubo(p0, X[i], Y[i]), g["Qc"][i])
        assert np.array_equal(M.disc_qubo(p0, X[i]), g["Qu"][i])
    c = M.disc_stats_faster_batch(g["Sc"], X, [[y] for y in Y], 1, 20, 6)
    u = M.disc_stats_faster_batch(g["Su"], X, None, 1, 20, 6)
    for nm, a, b in zip(["b_h", "b_o", "W_vh", "W_vo", "W_oo", "W_hh"], c, u):
        assert np.allclose(a, g[f"stat_c_{nm}"], rtol=0, atol=1e-14), nm
        assert np.allclose(b, g[f"stat_u_{nm}"], rtol=0, atol=1e-14), nm
    assert abs(M.disc_nll(g["Su"], Y) - float(g["nll"])) < 1e-6
    p1 = M.disc_train_step(p0, X, Y, g["Sc"], g["Su"], float(g["lr"]), "faster")
    for k, v in p1.items():
        assert np.allclose(v, g[f"w1_{k}"], rtol=0, atol=1e-14), k
    assert np.array_equal(p1["W_hh"], p0["W_hh"])          # Appendix B Q2: W_hh is never trained there


@pytest.mark.parametrize("name,one_hot", [("convdeep_binary.npz", False), ("convdeep_onehot.npz", True)])
def test_convdeep_matches_reference(name, one_hot):
    g = load(name)
    p0 = dict(kernel=g["w0_kernel"], W_seq=[g["w0_W_seq0"]], W_intra=[g["w0_W_intra0"]], W_hy=g["w0_W_hy"],
              W_oo=g["w0_W_oo"], b_conv=g["w0_b_conv"], b_seq=g["w0_b_seq"], b_out=g["w0_b_out"])
    n_lab = p0["b_out"].shape[0]
    X, Y = g["X"], g["Y"]
    for i in range(3):
        flat, pooled, patches = M.convdeep_context(X[i], p0["kernel"], 1, 2)
        lab = np.eye(n_lab)[Y[i]] if one_hot else np.array([float(Y[i])])
        assert np.allclose(M.convdeep_qubo(p0, flat, pooled, lab), g["Qc"][i], rtol=0, atol=1e-15)
        assert np.allclose(M.convdeep_qubo(p0, flat, pooled, None), g["Qu"][i], rtol=0, atol=1e-15)
        P = len(pooled)
        assert np.allclose(M.convdeep_probs(g["Su"][i], P + 12, one_hot), g["probs"][i])
        for tag, S, yy in (("c", g["Sc"][i], lab), ("u", g["Su"][i], None)):
            r = M.convdeep_stats(S.astype(np.float32), X[i], yy, P, [12], n_lab, patches)
            for nm, a in zip(["b_conv", "b_seq", "b_out", "kernel", "W_intra", "W_seq", "W_hy", "W_oo"], r):
                a = a[0] if isinstance(a, list) else a
                assert np.allclose(a, g[f"stat_{tag}_{nm}_{i}"], rtol=1e-6, atol=1e-7), (tag, nm)
    new, loss = M.convdeep_train_step(p0, X, Y, g["Sc"].astype(np.float32), g["Su"].astype(np.float32),
                                      float(g["lr"]), 1, 2, one_hot)
    assert abs(loss - float(g["loss"])) < 1e-6
    for k, ref in [("kernel", "kernel"), ("W_hy", "W_hy"), ("W_oo", "W_oo"), ("b_conv", "b_conv"), ("b_seq", "b_seq"),
                   ("b_out", "b_out")]:
        assert np.allclose(new[k], g[f"w1_{ref}"], rtol=1e-6, atol=1e-7), k
    assert np.allclose(new["W_seq"][0], g["w1_W_seq0"], rtol=1e-6, atol=1e-7)
    assert np.allclose(new["W_intra"][0], g["w1_W_intra0"], rtol=1e-6, atol=1e-7)


def test_rbm_matches_reference():
    g = load("rbm_discriminative.npz")
    W, U, bv, bh, bc = g["W0"], g["U0"], g["bv0"], g["bh0"], g["bc0"]
    x, y = g["x"], g["y"]
    onehot = np.eye(10, dtype=np.float32)[y[0]]
    assert np.allclose(M.rbm_sample_hidden(W, U, bh, x[0], onehot), g["ph0"], rtol=1e-5, atol=1e-6)
    assert np.allclose(M.rbm_sample_visible(W, bv, g["hbin"]), g["pv0"], rtol=1e-5, atol=1e-6)
    assert np.allclose(M.rbm_sample_class(U, bc, g["hbin"]), g["pc0"], rtol=1e-5, atol=1e-6)
    assert np.allclose(M.rbm_class_given_x(W, U, bh, bc, x[0]), g["pyx0"], rtol=1e-4, atol=1e-6)
    for s in range(2):
        new, probs, pred, err = M.rbm_discriminative_step(W, U, bv, bh, bc, x[s], y[s], float(g["lr"]))
        assert np.allclose(probs, g[f"probs{s}"], rtol=1e-4, atol=1e-6)
        assert np.array_equal(pred, g[f"pred{s}"])
        assert abs(err - float(g[f"err{s}"])) < 1e-5
        W, U, bv, bh, bc = new["W"], new["U"], new["b_v"], new["b_h"], new["b_c"]
        assert np.allclose(W, g[f"W{s + 1}"], rtol=1e-4, atol=1e-6)
        assert np.allclose(U, g[f"U{s + 1}"], rtol=1e-4, atol=1e-6)
        assert np.allclose(bh, g[f"bh{s + 1}"], rtol=1e-4, atol=1e-6)
        assert np.allclose(bc, g[f"bc{s + 1}"], rtol=1e-4, atol=1e-6)
        assert np.allclose(bv, g[f"bv{s + 1}"], rtol=1e-6, atol=1e-7)


def test_recorded_accuracy_by_exact_enumeration():
    """The reference's own recorded test accuracy (out/paper_data/.../e20_*_testacc_auc.pkl) is reproduced
    exactly by the ground state of the unclamped QUBO the oracle builds from the saved weights."""
    from sklearn.metrics import roc_auc_score
    g = load("pneumonia_h10_recorded_accuracy.npz")
    labels = g["labels"].astype(int)
    for k in range(3):
        off, diag = g[f"off_{k}"], g[f"diag_{k}"]
        n = diag.shape[1]
        X = ((np.arange(2 ** n)[:, None] >> np.arange(n)) & 1).astype(np.float64)
        quad = np.einsum("ri,ij,rj->r", X, off, X)
        pred = np.array([int(X[np.argmin(X @ d + quad), 0]) for d in diag])
        assert abs(np.mean(pred == labels) - float(g[f"acc_{k}"])) < 1e-12
        assert abs(roc_auc_score(labels, pred) - float(g[f"auc_{k}"])) < 1e-12


def _dg_close(a, ref, tag):
    assert np.allclose(M.array_digest(a), ref, rtol=1e-10, atol=1e-9), tag


def test_disc_qbm_at_the_c5_shapes_matches_reference():
    """Golden from discriminative_qbm.Disc_QBM at C5 (128 inputs, 10 one-hot labels, 512 hidden: n = 512 / 522), two images,
    100 reads x 1000 sweeps: initial draws, both QUBO builders, the loop statistics of both phases and the parameters after
    one training step of the reference equal the oracle's (large arrays compared through their digests)."""
    g = load("disc_qbm_loop_c5.npz")
    np.random.seed(77)                                           # Appendix B Q12
    p0 = M.disc_init_params(128, 10, 512, int(g["seed"]))
    for k, v in p0.items():
        _dg_close(v, g[f"w0_{k}_dg"], f"initial {k}")
    assert np.array_equal(p0["b_o"], g["w0_b_o"]) and np.array_equal(p0["W_oo"], g["w0_W_oo"])
    X, Y, Sc, Su = g["X"], g["Y"], g["Sc"], g["Su"]
    assert Sc.shape == (2, 100, 512) and Su.shape == (2, 100, 522)
    names = ["b_h", "b_o", "W_vh", "W_vo", "W_oo", "W_hh"]
    for i in range(2):
        _dg_close(M.disc_qubo(p0, X[i], Y[i]), g[f"Qc_dg_{i}"], f"Qc {i}")
        _dg_close(M.disc_qubo(p0, X[i]), g[f"Qu_dg_{i}"], f"Qu {i}")
        rc = M.disc_stats_loop(Sc[i], X[i], Y[i], 10, 128, 512)
        ru = M.disc_stats_loop(Su[i], X[i], None, 10, 128, 512)
        for nm, a, b in zip(names, rc, ru):
            _dg_close(a, g[f"stat_c_{nm}_dg_{i}"], f"clamped {nm} {i}")
            _dg_close(b, g[f"stat_u_{nm}_dg_{i}"], f"unclamped {nm} {i}")
    new = M.disc_train_step(p0, X, Y, Sc, Su, float(g["lr"]), "loop")
    for k, v in new.items():
        _dg_close(v, g[f"w1_{k}_dg"], f"after one step: {k}")
    assert np.allclose(new["b_o"], g["w1_b_o"], rtol=0, atol=1e-13) and np.allclose(new["W_oo"], g["w1_W_oo"], rtol=0, atol=1e-13)


def test_convdeep_at_the_c3_shapes_matches_reference():
    """Golden from Conv_Deep_QBM through src/train at C3 (18x18, 3x3 kernel, pool 2 -> 64 pooled, 128 sequential units, binary
    label: n = 192 / 193), two images, 1000 reads x 1000 sweeps."""
    g = load("convdeep_c3.npz")
    p0 = M.convdeep_init_params(64, [128], 1, 3, int(g["seed"]))
    flat = dict(kernel=p0["kernel"], W_seq0=p0["W_seq"][0], W_hy=p0["W_hy"], W_oo=p0["W_oo"], W_intra0=p0["W_intra"][0],
                b_conv=p0["b_conv"], b_seq=p0["b_seq"], b_out=p0["b_out"])
    for k, v in flat.items():
        _dg_close(v, g[f"w0_{k}_dg"], f"initial {k}")
    assert np.array_equal(p0["kernel"], g["w0_kernel"])
    X, Y, Sc, Su = g["X"], g["Y"], g["Sc"], g["Su"]
    assert Sc.shape == (2, 1000, 192) and Su.shape == (2, 1000, 193)
    names = ["b_conv", "b_seq", "b_out", "kernel", "W_intra", "W_seq", "W_hy", "W_oo"]
    for i in range(2):
        f, pooled, patches = M.convdeep_context(X[i], p0["kernel"], 1, 2)
        assert len(pooled) == 64
        lab = np.array([float(Y[i])])
        _dg_close(M.convdeep_qubo(p0, f, pooled, lab), g[f"Qc_dg_{i}"], f"Qc {i}")
        _dg_close(M.convdeep_qubo(p0, f, pooled, None), g[f"Qu_dg_{i}"], f"Qu {i}")
        for tag, S, yy in (("c", Sc[i], lab), ("u", Su[i], None)):
            r = M.convdeep_stats(S.astype(np.float32), X[i], yy, 64, [128], 1, patches)
            for nm, a in zip(names, r):
                a = a[0] if isinstance(a, list) else a
                assert np.allclose(M.array_digest(a), g[f"stat_{tag}_{nm}_dg_{i}"], rtol=1e-9, atol=1e-9), (tag, nm, i)
        assert np.allclose(M.convdeep_probs(Su[i].astype(np.float32), 192, False), g["probs"][i], rtol=1e-6)
    new, loss = M.convdeep_train_step(p0, X, Y, Sc.astype(np.float32), Su.astype(np.float32), float(g["lr"]), 1, 2, False)
    assert abs(loss - float(g["loss"])) < 1e-9
    flat1 = dict(kernel=new["kernel"], W_seq0=new["W_seq"][0], W_hy=new["W_hy"], W_oo=new["W_oo"], W_intra0=new["W_intra"][0],
                 b_conv=new["b_conv"], b_seq=new["b_seq"], b_out=new["b_out"])
    for k, v in flat1.items():
        assert np.allclose(M.array_digest(v), g[f"w1_{k}_dg"], rtol=1e-9, atol=1e-9), f"after one step: {k}"
    assert np.allclose(new["kernel"], g["w1_kernel"], rtol=1e-9, atol=1e-12)


def _majority_predictions_of_run(args):
    """Worker (forked process): Disc_QBM.predict for every test image of one recorded run through oracle.neal_sample."""
    from oracle import oracle as O
    off, diag, seed, sc = args
    pred = np.empty(len(diag), dtype=int)
    for i, d in enumerate(diag):
        smp, _ = O.neal_sample(off + np.diag(d), 30, 1000, seed=seed)
        if 0.2 < smp[:, 0].mean() < 0.8:
            smp, _ = O.neal_sample(off + np.diag(d), sc, 1000, seed=seed)
        pred[i] = int(np.round(smp[:, 0].mean()))                        # predict(): np.round(mean of the output column)
    return pred


def test_recorded_accuracy_through_the_neal_restatement_all_70_runs(oracle):
    """Pins oracle/neal_sa.c + the dimod/neal glue of oracle.py to data the reference holds: for all 70 PneumoniaMNIST
    last-epoch runs with h in {4,5,6,7,8,10,12} (out/paper_data/Pneumonia_param_doku/<h>_hnodes/_se*/)
    the unclamped QUBO of every one of the 624 test images is annealed by ``oracle.neal_sample`` -- the restated
    ``neal.SimulatedAnnealingSampler().sample(bqm, num_reads, num_sweeps=1000, seed=run seed)`` -- and the majority output
    bit over the reads (Disc_QBM.predict, faster_dqbm.py:1227-1241) must give exactly the (accuracy, AUC) the reference
    recorded for that run.  To keep the CPU suite short an image is first annealed with 30 reads; only when that vote is
    not decisive (output mean within (0.2, 0.8): near-degenerate ground states, a handful of images) it is re-annealed
    with the run's own sample count (50..600, from the run's parameter string).  The schedule is the runs' 1000 sweeps."""
    import multiprocessing as mp
    from sklearn.metrics import roc_auc_score
    g = load("pneumonia_last_epoch_recorded_accuracy.npz")
    labels = g["labels"].astype(int)
    runs = int(g["num_runs"])
    assert runs == 70
    jobs = [(g[f"off_{k}"], g[f"diag_{k}"].astype(np.float64), int(g[f"seed_{k}"]) % (2 ** 32), int(g[f"sc_{k}"]))
            for k in range(runs)]
    with mp.get_context("fork").Pool(min(os.cpu_count() or 4, 16)) as pool:
        preds = pool.map(_majority_predictions_of_run, jobs, chunksize=1)
    for k, pred in enumerate(preds):
        tag = (k, int(g[f"h_{k}"]), int(g[f"seed_{k}"]))
        assert abs(float(np.mean(pred == labels)) - float(g[f"acc_{k}"])) < 1e-12, tag
        assert abs(float(roc_auc_score(labels, pred)) - float(g[f"auc_{k}"])) < 1e-12, tag
