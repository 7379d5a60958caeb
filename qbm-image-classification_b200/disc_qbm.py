"""Batched, device-resident training step of the discriminative QBM.

Mirrors ``Disc_QBM`` (src/model/faster_dqbm.py -- the class qbm_main.py runs -- and its loop twin
src/model/discriminative_qbm.py): same constructor arguments, same parameter attribute names and
shapes, same ``create_qubo_matrix_from`` / ``get_samples`` / ``train_for_one_iteration`` / ``predict``
semantics, same learning rule ``param -= lr * (<.>_clamped - <.>_unclamped) / batch``.  What changes
is the execution: the per-image Python loop (two ``sampler.sample`` calls per image, faster_dqbm.py:
961-969) becomes one pass over the whole minibatch --

    X, Y --(GEMM, f64)--> 2B QUBOs --K0--> spin models --K1--> 2B x reads chains --K3--> moments
         --(small f64 GEMMs)--> parameter-shaped statistics --[all-reduce over ranks]--> SGD update

all on the current CUDA stream without a host synchronisation until the loss is read back.

``stats_mode`` selects which of the reference's two statistics functions is reproduced:
  "loop"    discriminative_qbm.py:696-760 (works for one-hot labels and any sizes; trains W_hh)
  "faster"  faster_dqbm.py:754-848 including its divergences (W_hh never
            trained, output-output term doubled, binary labels and n_hidden <= dim_input only)
"""
from __future__ import annotations

import os
import pickle
import random

import numpy as np
import torch

from . import _lib, dist as _d, ising, sampler as _s


def schedule_device(rng: torch.Tensor, num_sweeps: int):
    """Legacy neal beta range + geometric schedule from K0's ``range`` output ([B,2] = min non-zero
    |bias|, max total |bias|); all-zero problems get neal's [0.1, 1.0].  Returns (betas f32 [B,nb], spb)."""
    spb = int(max(1, num_sweeps // 1000.0))
    nb = -(-num_sweeps // spb) if num_sweeps > 0 else 0
    rng = rng.contiguous()
    B = rng.shape[0]
    betas = torch.empty((B, nb), dtype=torch.float32, device=rng.device)
    with torch.cuda.device(rng.device):
        rc = _lib.load().qbm_beta_schedule(rng.data_ptr(), B, nb, betas.data_ptr(), _s._stream_ptr(rng.device))
    _lib.check(rc)
    return betas, spb


class DiscQBM:
    def __init__(self, dim_input, num_classes, epochs=2, n_hidden_nodes=4, seed=77, solver="SA", restricted=False,
                 sample_count=20, anneal_steps=20, beta_eff=1, param_string="", load_path="", speicherort=None,
                 parallelize=False, use_one_hot_encoding=False, use_old_parallization=True,
                 stats_mode="faster", initial_states_generator="numpy", shared_stream=True, device=None,
                 process_group=None):
        if solver != "SA":
            raise ValueError("the B200 path implements solver='SA' only (D-Wave / BMS paths are out of scope)")
        if stats_mode not in ("loop", "faster"):
            raise ValueError("stats_mode must be 'loop' or 'faster'")
        self.epochs, self.seed = epochs, seed
        self.dim_input, self.n_hidden_nodes = int(dim_input), int(n_hidden_nodes)
        self.restricted, self.parallelize = bool(restricted), parallelize
        self.use_one_hot_encoding = use_one_hot_encoding
        self.n_output_nodes = int(num_classes) if use_one_hot_encoding else 1
        self.solver_string = solver
        self.sample_count, self.anneal_steps, self.beta_eff = int(sample_count), int(anneal_steps), beta_eff
        self.param_string, self.load_path, self.speicherort = param_string, load_path, speicherort
        self.stats_mode = stats_mode
        self.initial_states_generator = initial_states_generator
        self.shared_stream = shared_stream
        self.device = _s._require_cuda(device)
        self.pg = process_group
        if stats_mode == "faster" and (self.n_output_nodes != 1 or self.n_hidden_nodes > self.dim_input):
            # the reference's vectorised statistics raise for these shapes (Appendix B Q3)
            raise ValueError("stats_mode='faster' reproduces faster_dqbm.py, which only runs for binary labels and "
                             "n_hidden <= dim_input; use stats_mode='loop'")
        # ---- parameter initialisation: same draws, same order as faster_dqbm.py:77-83,192-223 ----
        h, no, di = self.n_hidden_nodes, self.n_output_nodes, self.dim_input
        W_hh = None if restricted else np.triu(np.random.uniform(-1, 1, (h, h)), k=1)   # drawn BEFORE the reseed (Q12)
        random.seed(seed)
        np.random.seed(seed)
        W_vh = np.random.uniform(-1, 1, (no + di, h))
        W_vo = np.random.uniform(-1, 1, (di, no))
        W_oo = np.triu(np.random.uniform(-1, 1, (no, no)), k=1)
        b_h = np.random.uniform(-1, 1, h)
        b_o = np.random.uniform(-1, 1, no)
        # all parameters live in ONE flat float64 device buffer (the layout of include/qbm_b200.h, K7-K9);
        # self._p holds views into it under the short names
        shapes = [("b_h", (h,)), ("b_o", (no,)), ("W_vh", (no + di, h)), ("W_vo", (di, no)), ("W_oo", (no, no))]
        if not restricted:
            shapes.append(("W_hh", (h, h)))
        total = sum(int(np.prod(sh)) for _, sh in shapes)
        assert total == _lib.load().qbm_disc_param_count(di, no, h, int(self.restricted))
        self._flat = torch.zeros(total, dtype=torch.float64, device=self.device)
        self._p = {"W_hh": None}
        pos = 0
        for nm, sh in shapes:
            cnt = int(np.prod(sh))
            self._p[nm] = self._flat[pos:pos + cnt].view(sh)
            pos += cnt
        self.set_params(W_vh=W_vh, W_vo=W_vo, W_oo=W_oo, b_h=b_h, b_o=b_o, W_hh=W_hh)
        self._init_cache = {}
        self.nll_per_batch = []
        self.step_count = 0
        self.keep_samples = False
        self.last_samples = None

    # ---- parameters: device tensors, exposed under the reference's attribute names -------------
    _NAMES = {"weights_all_visible_to_hidden": "W_vh", "weights_clamped_visible_to_output": "W_vo",
              "weights_output_output": "W_oo", "biases_hidden": "b_h", "biases_output": "b_o",
              "weights_hidden_hidden": "W_hh"}

    def set_params(self, **kw):
        for k, v in kw.items():
            if self._p.get(k) is None:
                if v is not None:
                    raise ValueError(f"{k} does not exist in a restricted model")
                continue
            v = torch.as_tensor(np.asarray(v, dtype=np.float64)).to(self.device)
            if v.shape != self._p[k].shape:
                raise ValueError(f"{k}: expected shape {tuple(self._p[k].shape)}, got {tuple(v.shape)}")
            self._p[k].copy_(v)

    def get_params(self) -> dict:
        return {k: (None if v is None else v.cpu().numpy()) for k, v in self._p.items()}

    def __getattr__(self, name):
        names = type(self)._NAMES
        if name in names and "_p" in self.__dict__:
            v = self._p[names[name]]
            return None if v is None else v.cpu().numpy()
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in type(self)._NAMES and "_p" in self.__dict__:
            self.set_params(**{type(self)._NAMES[name]: value})
        else:
            object.__setattr__(self, name, value)

    @property
    def weight_objects(self):
        p = self.get_params()
        return [p["W_vh"], p["W_vo"], p["b_h"], p["b_o"], p["W_oo"], p["W_hh"]]

    # ---- QUBO construction (faster_dqbm.py:225-284) ------------------------------------------------
    def _to_dev(self, a) -> torch.Tensor:
        """float64 device tensor from a numpy array (host -> device copy) or a tensor already on the device."""
        if torch.is_tensor(a):
            return a.to(self.device, torch.float64)
        return torch.as_tensor(np.asarray(a, dtype=np.float64)).to(self.device)

    def _labels(self, y_batch, B):
        return self._to_dev(y_batch).reshape(B, self.n_output_nodes)

    def build_qubos(self, X: torch.Tensor, Y: torch.Tensor | None) -> torch.Tensor:
        """Batched ``create_qubo_matrix_from`` (K7): float64 [B, n, n] on the device."""
        no, h, di = self.n_output_nodes, self.n_hidden_nodes, self.dim_input
        X = X.to(torch.float64).contiguous()
        B = X.shape[0]
        if X.shape != (B, di):
            raise ValueError(f"expected inputs of shape [B, {di}], got {tuple(X.shape)}")
        if Y is not None:
            Y = Y.to(torch.float64).contiguous()
            if Y.shape != (B, no):
                raise ValueError(f"expected labels of shape [{B}, {no}], got {tuple(Y.shape)}")
        n = h if Y is not None else no + h
        Q = torch.empty((B, n, n), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().qbm_disc_build_qubo(self._flat.data_ptr(), di, no, h, int(self.restricted), X.data_ptr(),
                                                 Y.data_ptr() if Y is not None else None, B, float(self.beta_eff),
                                                 Q.data_ptr(), _s._stream_ptr(self.device))
        _lib.check(rc)
        return Q

    def create_qubo_matrix_from(self, input_vector, label=None) -> np.ndarray:
        X = torch.as_tensor(np.asarray(input_vector, dtype=np.float64)).to(self.device)[None]
        Y = None if label is None else self._labels(np.atleast_1d(label), 1)
        return self.build_qubos(X, Y)[0].cpu().numpy()

    # ---- sampling ---------------------------------------------------------------------------------
    def _init_states(self, B, n):
        if self.initial_states_generator == "philox":
            return None
        key = (self.sample_count, n)
        if key not in self._init_cache:          # same RandomState(seed) draw for every call (Q6)
            self._init_cache[key] = torch.from_numpy(ising.initial_states_numpy(self.seed, self.sample_count, n)).to(self.device)
        return self._init_cache[key][None].expand(B, -1, -1).contiguous()

    def sample_batch(self, Q: torch.Tensor, first_image: int = 0) -> torch.Tensor:
        """int8 [B, sample_count, n] for a batch of QUBOs already on the device."""
        B, n, _ = Q.shape
        J, hh, _, rng = _s.qubo_to_ising_device(Q)
        betas, spb = schedule_device(rng, self.anneal_steps)
        flags = 2 if self.shared_stream else 0
        off = 0 if self.shared_stream else first_image * self.sample_count
        return _s.sa_sample(J, hh, betas, spb, self.sample_count, self.seed, chain_offset=off,
                            init_states=self._init_states(B, n), flags=flags).states

    def get_samples(self, input_vector, label=None) -> np.ndarray:
        """One image: int8 [sample_count, n] (``np.vstack`` of the reference's list of sample views)."""
        X = torch.as_tensor(np.asarray(input_vector, dtype=np.float64)).to(self.device)[None]
        Y = None if label is None else self._labels(np.atleast_1d(label), 1)
        return self.sample_batch(self.build_qubos(X, Y))[0].cpu().numpy()

    # ---- statistics -> parameter-shaped errors (sums over the local images), K8 -------------------
    def _errors(self, X, Y, mean_c, sec_c, mean_u, sec_u) -> torch.Tensor:
        """flat float64 [param_count + 1]: (clamped - unclamped) statistics in the parameter layout + the NLL sum."""
        no, h, di = self.n_output_nodes, self.n_hidden_nodes, self.dim_input
        B = X.shape[0]
        err = torch.empty(self._flat.numel() + 1, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().qbm_disc_errors(di, no, h, int(self.restricted), int(self.stats_mode == "faster"),
                                             X.data_ptr(), Y.data_ptr(), B, mean_c.data_ptr(),
                                             sec_c.data_ptr() if sec_c is not None else None, mean_u.data_ptr(),
                                             sec_u.data_ptr(), err.data_ptr(), _s._stream_ptr(self.device))
        _lib.check(rc)
        return err

    def train_for_one_iteration(self, x_batch, y_batch, learning_rate, nll=None, global_batch=None, first_image=0):
        """faster_dqbm.py:998-1064 / discriminative_qbm.py:875-951 for a whole minibatch.  With a
        process group, ``x_batch`` is this rank's shard, ``global_batch`` the minibatch size the errors
        are divided by and ``first_image`` the shard's offset.  Returns (errors_biases_output, avg loss)."""
        X = self._to_dev(x_batch).contiguous()
        B = X.shape[0]
        Y = self._labels(y_batch, B).contiguous()
        Sc = self.sample_batch(self.build_qubos(X, Y), first_image)
        Su = self.sample_batch(self.build_qubos(X, None), first_image)
        return self._step_from_samples(X, Y, Sc, Su, learning_rate, global_batch)

    def train_step_from_samples(self, x_batch, y_batch, samples_clamped, samples_unclamped, learning_rate, global_batch=None):
        """The same step from given sample sets (int8 0/1 CUDA tensors [B, R, h] and [B, R, n_out + h]) instead of the
        sampler's: statistics, errors, all-reduce and update as in ``train_for_one_iteration`` (what the golden tests use
        to run the reference's recorded sample sets through the kernels)."""
        X = self._to_dev(x_batch).contiguous()
        B = X.shape[0]
        Y = self._labels(y_batch, B).contiguous()
        h, no = self.n_hidden_nodes, self.n_output_nodes
        Sc, Su = samples_clamped.to(self.device).contiguous(), samples_unclamped.to(self.device).contiguous()
        if Sc.dtype != torch.int8 or Su.dtype != torch.int8 or Sc.shape[::2] != (B, h) or Su.shape[::2] != (B, no + h):
            raise ValueError("sample sets must be int8 tensors [B, R, h] (clamped) and [B, R, n_out + h] (unclamped)")
        return self._step_from_samples(X, Y, Sc, Su, learning_rate, global_batch)

    def _step_from_samples(self, X, Y, Sc, Su, learning_rate, global_batch):
        B = X.shape[0]
        mean_c, sec_c = _s.phase_stats(Sc, second=not self.restricted and self.stats_mode == "loop")
        mean_u, sec_u = _s.phase_stats(Su, second=True)
        if self.keep_samples:
            self.last_samples = (Sc, Su)
        flat = _d.all_reduce_sum_(self._errors(X, Y, mean_c, sec_c, mean_u, sec_u), self.pg)
        gb = float(global_batch if global_batch is not None else B)
        with torch.cuda.device(self.device):
            rc = _lib.load().qbm_sgd_apply(self._flat.data_ptr(), flat.data_ptr(), self._flat.numel(), float(learning_rate), gb,
                                           _s._stream_ptr(self.device))
        _lib.check(rc)
        h, no = self.n_hidden_nodes, self.n_output_nodes
        tail = torch.cat((flat[h:h + no] / gb, flat[-1:] / gb)).cpu().numpy()       # one device -> host read per step
        avg_loss = float(tail[-1])
        self.nll_per_batch.append(avg_loss)
        self.step_count += 1
        return tail[:no], avg_loss

    # ---- prediction (faster_dqbm.py:1227-1241) ------------------------------------------------------
    def predict_batch(self, X) -> np.ndarray:
        Xd = self._to_dev(X)
        Su = self.sample_batch(self.build_qubos(Xd, None))
        mean_u, _ = _s.phase_stats(Su, second=False)
        avg = mean_u[:, :self.n_output_nodes].cpu().numpy()
        if self.use_one_hot_encoding:
            return np.argmax(avg, axis=1)
        return np.round(avg).astype(int)[:, 0]

    def predict(self, data):
        S = self.get_samples(data)
        out = S[:, :self.n_output_nodes]
        avg = np.mean(out, axis=0)
        if self.use_one_hot_encoding:
            return int(np.argmax(avg)), out.tolist()
        return int(np.round(avg).astype(int)[0]), out.flatten()

    # ---- checkpoints: the reference's pickle format (faster_dqbm.py:1069-1077, 169-190) --------------------------
    def save_weights(self, title, path="out"):
        """pickle of ``weight_objects`` = [W_vh, W_vo, b_h, b_o, W_oo, W_hh] as float64 numpy arrays."""
        with open(f"{path}/{title}.pkl", "wb") as f:
            pickle.dump(self.weight_objects, f)

    def load_savepoint(self, savepoint):
        """Reads a weight pickle written by the reference or by :meth:`save_weights` (5 entries: semi-restricted
        runs without W_hh; 6 entries: fully connected)."""
        if not os.path.exists(savepoint):
            raise FileNotFoundError("Savepoint file not found")
        with open(savepoint, "rb") as f:
            loaded = pickle.load(f)
        assert len(loaded) in [5, 6]
        names = ["W_vh", "W_vo", "b_h", "b_o", "W_oo", "W_hh"][:len(loaded)]
        self.set_params(**{k: v for k, v in zip(names, loaded) if not (k == "W_hh" and (v is None or self.restricted))})

    # ---- epoch loop (faster_dqbm.py:1079-1166) ------------------------------------------------------------------
    def train_model(self, train_X, train_Y, val_X, val_Y, batch_size=8, learning_rate=0.005):
        """Same loop as the reference: minibatches in order (a shorter last batch included), weights pickled after
        every epoch when ``speicherort`` is set, validation accuracy per epoch -- with the validation set predicted in
        ONE batched launch instead of one sampler call per image.  Returns the history dict."""
        save_folder = None
        if self.speicherort is not None:
            save_folder = str(self.speicherort) + str(self.param_string)
            os.makedirs(save_folder, exist_ok=True)
        hist = {"errors_per_batch": [], "error_per_epoch": [], "nll_per_epoch": [], "acc_per_epoch": [], "auc_per_epoch": []}
        train_X, train_Y = np.asarray(train_X), np.asarray(train_Y)
        val_Y = np.asarray(val_Y)
        num_batches = max(1, len(train_X) // batch_size)
        for epoch in range(1, self.epochs + 1):
            epoch_errors, epoch_nll = 0.0, 0.0
            for b in range(0, len(train_X), batch_size):
                xb, yb = train_X[b:b + batch_size], train_Y[b:b + batch_size]
                if len(xb) == 0:
                    continue
                ebo, nll = self.train_for_one_iteration(xb, yb, learning_rate)
                hist["errors_per_batch"].append(float(np.mean(ebo)))
                epoch_errors += float(np.mean(ebo))
                epoch_nll += nll
            if save_folder is not None:
                self.save_weights(f"e{epoch}_{self.param_string}", save_folder)
            pred = self.predict_batch(val_X)
            truth = val_Y.argmax(axis=1) if val_Y.ndim == 2 and val_Y.shape[1] > 1 else val_Y.reshape(-1).astype(int)
            hist["acc_per_epoch"].append(float(np.mean(pred == truth)))
            try:
                from sklearn.metrics import roc_auc_score
                hist["auc_per_epoch"].append(float(roc_auc_score(truth, pred)) if len(np.unique(truth)) == 2 else float("nan"))
            except Exception:                                 # AUC is reporting only (src/metrics.py is out of scope)
                hist["auc_per_epoch"].append(float("nan"))
            hist["error_per_epoch"].append(epoch_errors / num_batches)
            hist["nll_per_epoch"].append(epoch_nll / num_batches)
        self.training_history = hist
        return hist
