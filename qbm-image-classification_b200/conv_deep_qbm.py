"""Batched, device-resident training step of the Conv-Deep QBM (the model ``cdqbm_main.py`` runs).

Mirrors ``Conv_Deep_QBM`` (src/model/cdqbm_state.py:13-215) and the functions that drive it --
``prepare_context`` (src/model/inference.py:16-44), ``build_unclamped_qubo`` / ``build_clamped_qubo``
(src/qubo/builder.py:21-110), ``run_unclamped`` / ``run_clamped`` (src/train/pipeline.py:13-36),
``train_one_iteration`` and ``get_average_configuration_single`` (src/train/train.py:12-253): same
constructor arguments, same parameter attribute names, shapes and initial draws, same variable layout
``[pooled conv | sequential layers | outputs]``, same learning rule.  What changes is the execution --
the reference's per-image loop (context, two QUBOs, two ``sampler.sample_Q`` calls, two statistics
passes per image) becomes one pass over the minibatch:

    X --K6--> feature maps, pooled units, patches --> 2B QUBOs --K0--> spin models --K1--> chains
      --K3--> moments --> parameter-shaped statistics --[all-reduce over ranks]--> SGD update

Supported configuration = what runs in the reference: deterministic pooling
(``pooling_size`` >= 2 or no pooling), ``hidden_bias_type`` "shared" or "none".  Probabilistic pooling
and per-unit conv biases crash inside the reference and are rejected here.
"""
from __future__ import annotations

import ctypes
import pickle
import random

import numpy as np
import torch

from . import _lib, dist as _d, ising, sampler as _s
from .disc_qbm import schedule_device


class ConvDeepQBM:
    def __init__(self, num_visible_nodes, num_lable_nodes, image_shape=(28, 28), seed=77, kernel_size=3, pooling_size=0,
                 pooling_type="deterministic", stride=1, sequential_layer_sizes=None, param_string="", load_path="",
                 speicherort=None, is_restricted=False, hidden_bias_type="none", solver="SA", anneal=1000, token="",
                 initial_states_generator="numpy", shared_stream=True, stats_dtype="float32", device=None,
                 process_group=None):
        if str(solver).upper() != "SA":
            raise ValueError("the B200 path implements solver='SA' only (the D-Wave adapter is out of scope)")
        if pooling_type != "deterministic":
            raise ValueError("only deterministic pooling is supported (probabilistic pooling raises inside the "
                             "reference's statistics, src/train/train.py:189)")
        if hidden_bias_type not in ("shared", "none"):
            raise ValueError("hidden_bias_type must be 'shared' or 'none' (per-unit conv biases need "
                             "model.pooled_units, which the reference never sets, src/train/train.py:176)")
        if stats_dtype not in ("float32", "float64"):
            raise ValueError("stats_dtype must be 'float32' (the reference's sample dtype) or 'float64'")
        self.kernel_size, self.pooling_size, self.pooling_type = int(kernel_size), int(pooling_size), pooling_type
        self.stride, self.image_shape = int(stride), tuple(int(v) for v in image_shape)
        self.sequential_layer_sizes = [int(s) for s in (sequential_layer_sizes or [])]
        self.num_visible, self.num_lable_nodes = num_visible_nodes, int(num_lable_nodes)
        self.seed, self.is_restricted, self.hidden_bias_type = seed, bool(is_restricted), hidden_bias_type
        self.param_string, self.load_path, self.speicherort = param_string, load_path, speicherort
        self.anneal = int(anneal)
        self.initial_states_generator, self.shared_stream, self.stats_dtype = initial_states_generator, shared_stream, stats_dtype
        self.device = _s._require_cuda(device)
        self.pg = process_group
        # ---- geometry (cdqbm_state.py:99-136, geometry.py:7-18,56-100) ----
        ih, iw = self.image_shape
        k, s = self.kernel_size, self.stride
        self.conv_layer_dim = ((ih - k) // s + 1, (iw - k) // s + 1)
        self.num_conv_units = self.conv_layer_dim[0] * self.conv_layer_dim[1]
        P = _lib.load().qbm_convdeep_num_pooled(ih, iw, k, s, self.pooling_size)
        if P < 1:
            raise ValueError(f"bad geometry: image {self.image_shape}, kernel {k}, stride {s}, pooling {pooling_size}")
        self.num_pooled_units = P
        self.num_hidden_units_per_layer = [self.num_conv_units] + self.sequential_layer_sizes
        self.num_active_units_per_layer = [self.num_conv_units, P] + self.sequential_layer_sizes
        self.num_active_units = sum(self.num_active_units_per_layer)
        self.num_hidden_nodes = self.num_conv_units + sum(self.sequential_layer_sizes)
        self.n_hidden = P + sum(self.sequential_layer_sizes)          # QUBO variables besides the outputs
        # ---- parameters: same draws in the same order as cdqbm_state.py:143-215 (MODEL.__init__ seeds first) ----
        np.random.seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        nl = self.num_lable_nodes
        kernel = np.random.uniform(-1, 1, (k, k))
        W_seq = [np.random.uniform(-1, 1, (self.num_active_units_per_layer[1 + i], n))
                 for i, n in enumerate(self.sequential_layer_sizes)]
        # np.triu of a 1-D draw: a size x size matrix whose rows repeat the vector (Appendix B Q10)
        W_intra = None if self.is_restricted else [np.triu(np.random.uniform(-1, 1, n)) for n in self.sequential_layer_sizes]
        W_hy = np.random.uniform(-1, 1, (self.num_active_units_per_layer[-1], nl))
        W_oo = np.triu(np.random.uniform(-1, 1, (nl, nl)), k=1)
        if hidden_bias_type == "shared":
            b_conv = np.random.uniform(-1, 1, 1)
        else:
            b_conv = np.zeros(self.sequential_layer_sizes)             # the reference's placeholder (:178), never used
        b_seq = np.random.uniform(-1, 1, sum(self.sequential_layer_sizes))
        b_out = np.random.uniform(-1, 1, nl)
        # trainable parameters live in ONE flat float64 device buffer (layout of include/qbm_b200.h, K10/K11);
        # self._p holds views into it.  With hidden_bias_type "none" the reference's b_conv placeholder is not a parameter.
        self._shared_bias = hidden_bias_type == "shared"
        sizes = self.sequential_layer_sizes
        widths = [P] + sizes
        shapes = ([("b_conv", (1,))] if self._shared_bias else []) + [("b_seq", (sum(sizes),)), ("b_out", (nl,)), ("kernel", (k, k))]
        shapes += [(("W_seq", li), (widths[li], widths[li + 1])) for li in range(len(sizes))]
        if not self.is_restricted:
            shapes += [(("W_intra", li), (sizes[li], sizes[li])) for li in range(len(sizes))]
        shapes += [("W_hy", (widths[-1], nl)), ("W_oo", (nl, nl))]
        self._sizes_c = (ctypes.c_int * max(1, len(sizes)))(*sizes)
        total = sum(int(np.prod(sh)) for _, sh in shapes)
        assert total == _lib.load().qbm_convdeep_param_count(P, len(sizes), self._sizes_c, nl, k, int(self.is_restricted),
                                                             int(self._shared_bias))
        self._flat = torch.zeros(total, dtype=torch.float64, device=self.device)
        self._p = {"W_seq": [None] * len(sizes), "W_intra": None if self.is_restricted else [None] * len(sizes),
                   "b_conv": None}
        pos = 0
        for nm, sh in shapes:
            cnt = int(np.prod(sh))
            view = self._flat[pos:pos + cnt].view(sh)
            if isinstance(nm, tuple):
                self._p[nm[0]][nm[1]] = view
            else:
                self._p[nm] = view
            pos += cnt
        self._b_conv_placeholder = None if self._shared_bias else np.asarray(b_conv, dtype=np.float64)
        self.set_params(kernel=kernel, W_seq=W_seq, W_intra=W_intra, W_hy=W_hy, W_oo=W_oo, b_seq=b_seq, b_out=b_out,
                        **({"b_conv": b_conv} if self._shared_bias else {}))
        self._init_cache = {}
        self.keep_samples = False
        self.last_samples = None
        self.step_count = 0

    # ---- parameters: device tensors, exposed under the reference's attribute names -------------
    _NAMES = {"kernel_weights": "kernel", "weights_sequential_layer": "W_seq", "weights_hidden_to_output": "W_hy",
              "weights_output_output": "W_oo", "weights_interlayer_sequential": "W_intra", "biases_conv_units": "b_conv",
              "biases_sequential_units": "b_seq", "biases_output": "b_out"}

    def _copy_in(self, dst, v, name):
        v = torch.as_tensor(np.asarray(v, dtype=np.float64)).to(self.device)
        if v.shape != dst.shape:
            raise ValueError(f"{name}: expected shape {tuple(dst.shape)}, got {tuple(v.shape)}")
        dst.copy_(v)

    def set_params(self, **kw):
        for k, v in kw.items():
            if k == "b_conv" and not self._shared_bias:
                self._b_conv_placeholder = np.asarray(v, dtype=np.float64)
            elif self._p.get(k) is None:
                if v is not None:
                    raise ValueError(f"{k} does not exist in this model (restricted / no sequential layers)")
            elif isinstance(self._p[k], list):
                if len(v) != len(self._p[k]):
                    raise ValueError(f"{k}: expected {len(self._p[k])} layers")
                for dst, a in zip(self._p[k], v):
                    self._copy_in(dst, a, k)
            else:
                self._copy_in(self._p[k], v, k)

    def get_params(self) -> dict:
        out = {}
        for k, v in self._p.items():
            out[k] = None if v is None else ([a.cpu().numpy() for a in v] if isinstance(v, list) else v.cpu().numpy())
        if not self._shared_bias:
            out["b_conv"] = self._b_conv_placeholder.copy()
        return out

    def __getattr__(self, name):
        names = type(self)._NAMES
        if name in names and "_p" in self.__dict__:
            return self.get_params()[names[name]]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in type(self)._NAMES and "_p" in self.__dict__:
            self.set_params(**{type(self)._NAMES[name]: value})
        else:
            object.__setattr__(self, name, value)

    @property
    def weight_objects(self):
        p = self.get_params()
        return [p["kernel"], p["W_seq"], p["W_hy"], p["W_oo"], p["W_intra"], p["b_conv"], p["b_seq"], p["b_out"]]

    # ---- context (inference.py:16-44) ---------------------------------------------------------------
    def prepare_context_batch(self, X):
        """(fmap f64 [B, oh*ow], pooled_idx int32 [B, P], patches f64 [B, P, k, k]) for images [B, ih, iw] (K6)."""
        Xd = torch.as_tensor(np.asarray(X) if not torch.is_tensor(X) else X).to(self.device, torch.float64).contiguous()
        if Xd.dim() == 2:
            Xd = Xd[None]
        B, ih, iw = Xd.shape
        if (ih, iw) != self.image_shape:
            raise ValueError(f"images are {ih}x{iw}, the model was built for {self.image_shape}")
        k, P = self.kernel_size, self.num_pooled_units
        fmap = torch.empty((B, self.num_conv_units), dtype=torch.float64, device=self.device)
        pooled = torch.empty((B, P), dtype=torch.int32, device=self.device)
        patches = torch.empty((B, P, k, k), dtype=torch.float64, device=self.device)
        L = _lib.load()
        with torch.cuda.device(self.device):
            rc = L.qbm_convdeep_context(Xd.data_ptr(), self._p["kernel"].data_ptr(), B, ih, iw, k, self.stride,
                                        self.pooling_size, fmap.data_ptr(), pooled.data_ptr(), patches.data_ptr(),
                                        _s._stream_ptr(self.device))
        _lib.check(rc)
        return fmap, pooled, patches

    # ---- QUBO construction (builder.py:21-110) -------------------------------------------------------
    def _struct(self):
        return (self.num_pooled_units, len(self.sequential_layer_sizes), self._sizes_c, self.num_lable_nodes, self.kernel_size,
                int(self.is_restricted), int(self._shared_bias))

    def build_qubos(self, fmap: torch.Tensor, pooled: torch.Tensor, Y: torch.Tensor | None, beta_eff: float = 1.0):
        """float64 [B, n, n] (K10): n = n_hidden (+ num_lable_nodes when ``Y`` is None, i.e. unclamped)."""
        nh, nl = self.n_hidden, self.num_lable_nodes
        B = fmap.shape[0]
        n = nh + (nl if Y is None else 0)
        fmap, pooled = fmap.contiguous(), pooled.contiguous()
        if Y is not None:
            Y = Y.to(torch.float64).contiguous()
        Q = torch.empty((B, n, n), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().qbm_convdeep_build_qubo(self._flat.data_ptr(), *self._struct(), fmap.data_ptr(), fmap.shape[1],
                                                     pooled.data_ptr(), Y.data_ptr() if Y is not None else None, B,
                                                     float(beta_eff), Q.data_ptr(), _s._stream_ptr(self.device))
        _lib.check(rc)
        return Q

    # ---- sampling (pipeline.py:20,35 -> sampler.py:26-33) -----------------------------------------------
    def _init_states(self, B, n, num_reads):
        if self.initial_states_generator == "philox":
            return None
        key = (num_reads, n)
        if key not in self._init_cache:              # LocalSASampler passes the same seed on every call (Q6)
            self._init_cache[key] = torch.from_numpy(ising.initial_states_numpy(self.seed, num_reads, n)).to(self.device)
        return self._init_cache[key][None].expand(B, -1, -1).contiguous()

    def sample_batch(self, Q: torch.Tensor, num_reads: int, first_image: int = 0) -> torch.Tensor:
        B, n, _ = Q.shape
        J, hh, _, rng = _s.qubo_to_ising_device(Q)
        betas, spb = schedule_device(rng, self.anneal)
        flags = 2 if self.shared_stream else 0
        off = 0 if self.shared_stream else first_image * int(num_reads)
        return _s.sa_sample(J, hh, betas, spb, int(num_reads), self.seed, chain_offset=off,
                            init_states=self._init_states(B, n, int(num_reads)), flags=flags).states

    # ---- class probabilities (pipeline.py:22-28) ------------------------------------------------------
    def _probs(self, mean_u: torch.Tensor, one_hot: bool) -> torch.Tensor:
        out = mean_u[:, self.n_hidden:].to(torch.float32)
        if not one_hot:
            p1 = out[:, 0].to(torch.float64).clamp(1e-12, 1 - 1e-12)
            return torch.stack((1.0 - p1, p1), dim=1).to(torch.float32)
        s = out.sum(dim=1, keepdim=True)
        uni = torch.full_like(out, 1.0 / out.shape[1])
        return torch.where(s > 0, out / torch.where(s > 0, s, torch.ones_like(s)), uni)

    # ---- statistics -> parameter-shaped errors + loss, summed over the local images (train.py:135-253), K11 ------
    def _errors(self, patches, Ylab, y, one_hot, mc, sc, mu, su) -> torch.Tensor:
        B = patches.shape[0]
        err = torch.empty(self._flat.numel() + 1, dtype=torch.float64, device=self.device)
        y32 = y.to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            rc = _lib.load().qbm_convdeep_errors(*self._struct(), int(self.stats_dtype == "float32"), int(bool(one_hot)),
                                                 patches.data_ptr(), Ylab.data_ptr(), y32.data_ptr(), B, mc.data_ptr(),
                                                 sc.data_ptr(), mu.data_ptr(), su.data_ptr(), err.data_ptr(),
                                                 _s._stream_ptr(self.device))
        _lib.check(rc)
        return err

    def _labels(self, Y, B, one_hot):
        y = (Y if torch.is_tensor(Y) else torch.as_tensor(np.asarray(Y))).to(self.device).reshape(B).to(torch.int64)
        if one_hot:
            return y, torch.nn.functional.one_hot(y, self.num_lable_nodes).to(torch.float64)
        return y, y.to(torch.float64)[:, None]

    def train_one_iteration(self, X, Y, num_reads: int, beta_eff: float, lr: float, one_hot: bool = False,
                            global_batch=None, first_image: int = 0) -> float:
        """src/train/train.py:12-132 for a whole minibatch; returns the mean loss.  With a process group,
        ``X`` is this rank's shard, ``global_batch`` the minibatch size and ``first_image`` the shard's offset."""
        fmap, pooled, patches = self.prepare_context_batch(X)
        B = fmap.shape[0]
        y, Ylab = self._labels(Y, B, one_hot)
        Ylab = Ylab.contiguous()
        Sc = self.sample_batch(self.build_qubos(fmap, pooled, Ylab, beta_eff), num_reads, first_image)
        Su = self.sample_batch(self.build_qubos(fmap, pooled, None, beta_eff), num_reads, first_image)
        return self._step_from_samples(patches, Ylab, y, one_hot, Sc, Su, lr, global_batch)

    def train_step_from_samples(self, X, Y, samples_clamped, samples_unclamped, lr: float, one_hot: bool = False,
                                global_batch=None) -> float:
        """The same step from given sample sets (int8 0/1 CUDA tensors [B, R, n_hidden] and [B, R, n_hidden + labels])
        instead of the sampler's (the golden tests run the reference's recorded sample sets through the kernels)."""
        fmap, pooled, patches = self.prepare_context_batch(X)
        y, Ylab = self._labels(Y, fmap.shape[0], one_hot)
        Sc, Su = samples_clamped.to(self.device).contiguous(), samples_unclamped.to(self.device).contiguous()
        if Sc.dtype != torch.int8 or Su.dtype != torch.int8:
            raise ValueError("sample sets must be int8 tensors")
        return self._step_from_samples(patches, Ylab.contiguous(), y, one_hot, Sc, Su, lr, global_batch)

    def _step_from_samples(self, patches, Ylab, y, one_hot, Sc, Su, lr, global_batch) -> float:
        B = patches.shape[0]
        mc, sc = _s.phase_stats(Sc)
        mu, su = _s.phase_stats(Su)
        if self.keep_samples:
            self.last_samples = (Sc, Su)
        flat = _d.all_reduce_sum_(self._errors(patches, Ylab, y, one_hot, mc, sc, mu, su), self.pg)
        gb = float(global_batch if global_batch is not None else B)
        with torch.cuda.device(self.device):
            rc = _lib.load().qbm_sgd_apply(self._flat.data_ptr(), flat.data_ptr(), self._flat.numel(), float(lr), gb,
                                           _s._stream_ptr(self.device))
        _lib.check(rc)
        self.step_count += 1
        return float(flat[-1].item() / max(1.0, gb))

    # ---- prediction (cdqbm_main.py:119-127 -> run_unclamped) --------------------------------------------
    def predict_proba_batch(self, X, num_reads: int, beta_eff: float = 1.0, one_hot: bool = False) -> np.ndarray:
        fmap, pooled, _ = self.prepare_context_batch(X)
        Su = self.sample_batch(self.build_qubos(fmap, pooled, None, beta_eff), num_reads)
        mu, _ = _s.phase_stats(Su, second=False)
        return self._probs(mu, one_hot).cpu().numpy()

    # ---- epoch loop (src/train/train.py:256-289) and checkpoints (src/model/model_ab.py:33-35) ----------------------
    def train_model(self, train_x, train_y, batch_size, epochs, lr, sample_count, beta_eff, one_hot: bool = False):
        """``train_model(model, ...)`` of the reference as a method: returns the running average loss after every
        minibatch (``epoch_loss_list``)."""
        n = len(train_x)
        epoch_loss_list = []
        for _ in range(1, epochs + 1):
            epoch_loss = 0.0
            for idx, b in enumerate(range(0, n, batch_size)):
                loss = self.train_one_iteration(train_x[b:b + batch_size], train_y[b:b + batch_size], sample_count, beta_eff, lr,
                                                one_hot=one_hot)
                epoch_loss += loss
                epoch_loss_list.append(epoch_loss / (idx + 1))
        return epoch_loss_list

    def save_weights(self, title, path=""):
        with open(f"{path}/{title}.pkl", "wb") as f:
            pickle.dump(self.weight_objects, f)


def train_model(model: ConvDeepQBM, train_x, train_y, batch_size, epochs, lr, sample_count, beta_eff, one_hot: bool = False):
    """Call-compatible with ``src/train/train.py::train_model`` for a :class:`ConvDeepQBM`."""
    return model.train_model(train_x, train_y, batch_size, epochs, lr, sample_count, beta_eff, one_hot)
