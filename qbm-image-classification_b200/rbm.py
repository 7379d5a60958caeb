"""Boundary B3: ``ClassificationRBM`` with the reference's attribute names,
shapes and method signatures (src/ClassificationRBM.py:12-157), executed by the tcgen05 GEMM and the
fused elementwise kernels of libqbm_b200.so.

Reference attributes ``weights [V,H]``, ``visible_bias [V]``, ``hidden_bias [H]``, ``class_weights [C,H]``,
``class_bias [C]`` are float32 CUDA tensors (views into 16-byte-row padded storage).  Methods:

* ``sample_hidden(v, y_onehot)`` / ``sample_visible(h)`` / ``sample_class(h)``   (:43-60)
* ``sample_class_given_x(x)``                                                   (:62-86)
* ``discriminative_training(x, y, factor=1) -> (error, predicted, class_probabilities)``  (:101-146)
* ``cd1_training(v0, y0)`` -- the CD-1 step assembled from the three primitives (the reference keeps
  ``k`` and the primitives but never wires them)
"""
from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np
import torch

from . import _lib
from .sampler import _require_cuda, _stream_ptr


# models that hold captured data-parallel steps: NCCL cannot destroy a communicator while CUDA graphs that captured its
# collectives are alive, so torch.distributed.destroy_process_group is wrapped (once) to release them first
_DP_GRAPH_MODELS = weakref.WeakSet()


def _hook_destroy_process_group():
    import torch.distributed as dist
    if getattr(dist.destroy_process_group, "_qbm_b200_hook", False):
        return
    orig = dist.destroy_process_group

    def destroy_process_group(group=None):
        # every rank runs this wrapper (SPMD) whatever models it still holds, so the ranks meet HERE, once -- a peer may still
        # be reading this rank's last gradient -- and the models are then released without further collectives
        try:
            if dist.is_initialized() and (group is None or group is dist.group.WORLD):
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                dist.barrier()
        except Exception:
            pass
        for m in list(_DP_GRAPH_MODELS):
            m._graphs.clear()
            m._close_peer(collective=False)
        return orig(group)

    destroy_process_group._qbm_b200_hook = True
    dist.destroy_process_group = destroy_process_group


def _ld4(c: int) -> int:
    return (c + 3) & ~3


def _padded(rows: int, cols: int, dev) -> torch.Tensor:
    return torch.zeros((rows, _ld4(cols)), dtype=torch.float32, device=dev)


def gemm_tf32(A: torch.Tensor, B: torch.Tensor, alpha=1.0, beta=0.0, Cin=None, bias=None, act=0, want_transposed=False):
    """C = act(alpha * A @ B.T + bias) + beta * Cin on the tcgen05 TF32 pipeline (A [M,K], B [N,K], K % 4 == 0)."""
    L = _lib.load()
    assert A.is_cuda and B.is_cuda and A.dtype == torch.float32 and B.dtype == torch.float32
    A = A.contiguous(); B = B.contiguous()
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    Ct = torch.empty((N, M), dtype=torch.float32, device=A.device) if want_transposed else None
    if Cin is not None:
        Cin = Cin.contiguous()
    with torch.cuda.device(A.device):
        rc = L.qbm_gemm_tf32(A.data_ptr(), K, B.data_ptr(), K, M, N, K, float(alpha), float(beta),
                             Cin.data_ptr() if Cin is not None else None, N, bias.data_ptr() if bias is not None else None,
                             int(act), C.data_ptr(), N, Ct.data_ptr() if Ct is not None else None, M, _stream_ptr(A.device))
    _lib.check(rc)
    return (C, Ct) if want_transposed else C


class B200ClassificationRBM:
    def __init__(self, num_visible, num_hidden, k, num_classes=2, learning_rate=0.05, sparse_constant=0.00,
                 use_cuda=True, seed=42, device=None, process_group=None, use_graphs=True, peer_reduce=True):
        # same RNG protocol as the reference (:14-15, :26-30): CPU generators, then moved to the device
        np.random.seed(seed)
        torch.manual_seed(seed)
        self.seed = seed
        self.num_visible, self.num_hidden, self.k = int(num_visible), int(num_hidden), k
        self.learning_rate = learning_rate
        self.use_cuda = True
        self.num_classes = int(num_classes)
        self.sparse_constant = sparse_constant
        if self.num_classes > 32:
            raise ValueError("the fused class kernels support at most 32 classes")
        self.device = _require_cuda(device)
        self.pg = process_group            # data-parallel minibatches: deltas are all-reduced
        V, H, C = self.num_visible, self.num_hidden, self.num_classes
        w = torch.randn(V, H) * 0.1
        self._W = _padded(V, H, self.device); self._W[:, :H] = w.to(self.device)
        self._Wt = _padded(H, V, self.device); self._Wt[:, :V] = w.t().to(self.device)
        self._U = _padded(C, H, self.device)
        self.visible_bias = (torch.ones(V) * 0.5).to(self.device)
        self.hidden_bias = torch.zeros(H, device=self.device)
        self.class_bias = torch.zeros(C, device=self.device)
        self._ws = None
        self._grad = None
        self._step = 0
        # a training step is a fixed sequence of 4-9 small launches (plus the NCCL all-reduce of the data-parallel mode):
        # from the second step of a given shape on it is replayed as one CUDA graph (static input / output buffers; the
        # CD-1 step counter lives on the device)
        self.use_graphs = bool(use_graphs)
        self._graphs = {}
        self._ctr = None                   # device counters of the captured steps: [step * world (Philox streams), step (tokens)]
        self._ctr_val = -1
        # data-parallel mode: the gradient all-reduce and the update as one pass over NVLink peer memory (csrc/peer.cu)
        # when every rank of the group can map the others' gradient buffers (one node); else NCCL all-reduce + apply
        self.peer_reduce = bool(peer_reduce) and os.environ.get("QBM_RBM_PEER_REDUCE", "1") != "0"     # (env: A/B measurements)
        self._peer = None                  # dict(own=ptr, bases=ctypes array, imported=[ptrs]) once set up; False = not possible
        self.acc_per_epoch_list = []
        self.auc_per_epoch_list = []

    # ---- reference attribute names ---------------------------------------------------------------
    @property
    def weights(self):
        return self._W[:, :self.num_hidden]

    @weights.setter
    def weights(self, w):
        # the getter returns a view of the padded storage, so `m.weights += d` (the reference's own update_weights
        # pattern, ClassificationRBM.py:88-99) hands that view back: copy before touching the storage
        w = torch.as_tensor(w, dtype=torch.float32).to(self.device).clone()
        if w.shape != (self.num_visible, self.num_hidden):
            raise ValueError(f"weights must be [{self.num_visible}, {self.num_hidden}], got {tuple(w.shape)}")
        self._W.zero_(); self._W[:, :self.num_hidden] = w
        self.sync_transposed_weights()

    def sync_transposed_weights(self):
        """Rebuild the K-major copy W^T the GEMMs read from W (needed after in-place edits through the ``weights`` view,
        e.g. ``m.weights.add_(d)``; the setter and the step functions do it themselves)."""
        self._Wt.zero_(); self._Wt[:, :self.num_visible] = self._W[:, :self.num_hidden].t()

    @property
    def class_weights(self):
        return self._U[:, :self.num_hidden]

    @class_weights.setter
    def class_weights(self, u):
        u = torch.as_tensor(u, dtype=torch.float32).to(self.device).clone()
        if u.shape != (self.num_classes, self.num_hidden):
            raise ValueError(f"class_weights must be [{self.num_classes}, {self.num_hidden}], got {tuple(u.shape)}")
        self._U.zero_(); self._U[:, :self.num_hidden] = u

    # ---- helpers ------------------------------------------------------------------------------------
    def _workspace(self, B):
        """Caller-owned workspace of the step functions; grows, never shrinks (captured steps point into it, so a new
        allocation drops them)."""
        L = _lib.load()
        nbytes = L.qbm_rbm_workspace_bytes(B, self.num_visible, self.num_hidden, self.num_classes)
        if self._ws is None or self._ws.numel() * 4 < nbytes:
            self._graphs.clear()
            self._ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
        return self._ws

    def _pad_rows(self, x, cols):
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device)
        if x.dim() != 2 or x.shape[1] != cols:
            raise ValueError(f"expected a [B, {cols}] matrix, got {tuple(x.shape)}")
        if cols == _ld4(cols) and x.is_contiguous():
            return x
        out = _padded(x.shape[0], cols, self.device)
        out[:, :cols] = x
        return out

    def _labels(self, y, B):
        y = torch.as_tensor(y).to(self.device)
        if y.dim() == 2:                       # one-hot rows, as sample_hidden receives them
            y = y.argmax(dim=1)
        if y.shape != (B,):
            raise ValueError(f"expected {B} labels, got shape {tuple(y.shape)}")
        return y.to(torch.int32).contiguous()

    def _call(self, fn, *args):
        with torch.cuda.device(self.device):
            rc = fn(*args, _stream_ptr(self.device))
        _lib.check(rc)

    # ---- data-parallel minibatches ---------------------------------------------------------------------
    def _world(self):
        if self.pg is None:
            return 1
        import torch.distributed as dist
        return dist.get_world_size(self.pg)

    def _rank(self):
        if self.pg is None:
            return 0
        import torch.distributed as dist
        return dist.get_rank(self.pg)

    def _grad_buffer(self):
        """The flat gradient buffer of the data-parallel steps ([dW | dU | db_v | db_h | db_c | loss sum], include/qbm_b200.h):
        written by the gradient kernels, all-reduced as it is, consumed by the fused apply."""
        if getattr(self, "_grad", None) is None:
            n = _lib.load().qbm_rbm_grad_count(self.num_visible, self.num_hidden, self.num_classes)
            self._grad = torch.zeros(n, dtype=torch.float32, device=self.device)
        return self._grad

    # ---- CUDA-graph replay of the training steps -----------------------------------------------------------------
    def _graph_entry(self, key, B, launch):
        """None on the first step of a key (the caller runs it eagerly, which also warms the kernels up), afterwards the
        captured step: static buffers x [B, ld4(V)], y int32 [B], out = [probs (B x ld4(C)) | loss] and pred int32 [B]."""
        self._workspace(B)                      # (a new workspace drops every captured step)
        # the captured launches hold raw pointers: a parameter tensor that was re-assigned (not updated in place) gets a new capture
        key = key + tuple(t.data_ptr() for t in (self._W, self._Wt, self._U, self.visible_bias, self.hidden_bias, self.class_bias))
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 32:         # shapes / learning rates / re-assigned tensors keep changing: start over
                self._graphs.clear()
            self._graphs[key] = "warm"
            return None
        if ent == "warm":
            V, C = self.num_visible, self.num_classes
            ent = {"x": _padded(B, V, self.device), "y": torch.zeros(B, dtype=torch.int32, device=self.device),
                   "out": torch.zeros(B * _ld4(C) + 4, dtype=torch.float32, device=self.device),
                   "pred": torch.zeros(B, dtype=torch.int32, device=self.device)}
            if self.pg is not None:
                _DP_GRAPH_MODELS.add(self)
                _hook_destroy_process_group()
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.device(self.device), torch.cuda.graph(g, capture_error_mode="thread_local"):
                launch(ent)
            ent["graph"] = g
            self._graphs[key] = ent
        return ent

    def release_graphs(self):
        """Drop every captured step (re-captured on demand) and, in data-parallel mode, the peer-mapped gradient buffers (set up
        again on demand; a COLLECTIVE call there: the ranks meet before the memory goes).  NCCL cannot tear a communicator
        down while CUDA graphs that captured its collectives are alive either: ``torch.distributed.destroy_process_group`` is
        wrapped to call this for every live data-parallel model; call it yourself if you destroy the group some other way."""
        self._graphs.clear()
        broken = self.peer_error()
        self._close_peer()
        if broken:
            raise RuntimeError("a rank of the process group never delivered its gradient (waited ~2 min): the data-parallel "
                               "updates since then were skipped on this rank")

    # ---- peer-memory gradient buffers (data-parallel mode) -----------------------------------------------------------
    def _peer_setup(self):
        """Allocate this rank's IPC-shared gradient allocation, exchange the handles over the process group and map the
        peers'.  Every rank ends with the same answer (an all-reduce of the outcome), so the ranks never mix the two paths."""
        if self._peer is not None:
            return self._peer
        import torch.distributed as dist
        L = _lib.load()
        world, rank = self._world(), self._rank()
        V, H, C = self.num_visible, self.num_hidden, self.num_classes
        own, imported, ok = ctypes.c_void_p(), [], 1
        handles = [None] * world
        try:
            if not self.peer_reduce or world > 16:
                raise RuntimeError("peer reduce switched off")
            with torch.cuda.device(self.device):
                _lib.check(L.qbm_peer_alloc(L.qbm_rbm_peer_bytes(V, H, C), ctypes.byref(own)))
                h = (ctypes.c_ubyte * 64)()
                _lib.check(L.qbm_peer_export(own, h))
            mine = bytes(h)
        except Exception:
            ok, mine = 0, b""
        dist.all_gather_object(handles, mine, group=self.pg)
        bases = (ctypes.c_void_p * world)()
        if ok and all(len(x) == 64 for x in handles):
            try:
                with torch.cuda.device(self.device):
                    for p in range(world):
                        if p == rank:
                            bases[p] = own.value
                        else:
                            q = ctypes.c_void_p()
                            _lib.check(L.qbm_peer_import((ctypes.c_ubyte * 64).from_buffer_copy(handles[p]), ctypes.byref(q)))
                            imported.append(q)
                            bases[p] = q.value
            except Exception:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)         # (also: every rank has mapped before anyone signals)
        if int(flag.item()) == 1:
            self._peer = {"own": own, "bases": bases, "imported": imported,
                          "count": int(L.qbm_rbm_grad_count(V, H, C))}
            _DP_GRAPH_MODELS.add(self)
            _hook_destroy_process_group()
        else:
            self._peer = {"own": own, "bases": None, "imported": imported, "count": 0}
            self._close_peer(collective=False)
            self._peer = False
        return self._peer

    def _close_peer(self, collective=True):
        """Unmap the peers' allocations and free this rank's.  A peer may still be reading this rank's last gradient, so the
        ranks meet first (device synchronised, then a barrier over the group): in data-parallel mode `release_graphs` is a
        collective call."""
        pr = self._peer
        if not pr:
            return
        self._peer = None
        L = _lib.load()
        try:
            torch.cuda.synchronize(self.device)
            if collective and pr.get("bases") is not None:
                import torch.distributed as dist
                if dist.is_initialized():
                    dist.barrier(group=self.pg)
            with torch.cuda.device(self.device):
                for q in pr["imported"]:
                    L.qbm_peer_close(q)
                if pr["own"].value:
                    L.qbm_peer_free(pr["own"])
        except Exception:
            pass

    def peer_error(self) -> bool:
        """True when a wait for the peers' gradients ever timed out on this rank (a peer died or fell out of step)."""
        if not self._peer:
            return False
        out = ctypes.c_uint(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().qbm_rbm_peer_error(self._peer["own"], self.num_visible, self.num_hidden, self.num_classes,
                                                      ctypes.byref(out)))
        return out.value != 0

    def __del__(self):
        try:
            self._graphs.clear()
            self._close_peer(collective=False)
        except Exception:
            pass

    def _dp_reduce_apply(self, launch_grads, scale, loss_p, gb, token_host, tick_dev_p):
        """Data-parallel tail of a step: `launch_grads(grad_ptr)` writes the shard's gradient sums, then either the peer-memory
        pass (signal, wait, reduce in rank order, apply) or NCCL all-reduce + apply."""
        L = _lib.load()
        V, H, C = self.num_visible, self.num_hidden, self.num_classes
        params = (self._W.data_ptr(), self._Wt.data_ptr(), self._U.data_ptr(), self.visible_bias.data_ptr(),
                  self.hidden_bias.data_ptr(), self.class_bias.data_ptr())
        pr = self._peer_setup()
        if pr:
            parity = self._step & 1
            launch_grads(pr["own"].value + parity * pr["count"] * 4)
            self._call(L.qbm_rbm_apply_grad_peer, *params, pr["bases"], self._world(), self._rank(), parity, V, H, C,
                       float(scale), float(self.sparse_constant), loss_p, float(1.0 / gb), ctypes.c_uint(token_host), tick_dev_p)
            return
        import torch.distributed as dist
        grad = self._grad_buffer()
        launch_grads(grad.data_ptr())
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.pg)
        self._call(L.qbm_rbm_apply_grad, *params, grad.data_ptr(), V, H, C, float(scale), float(self.sparse_constant), loss_p,
                   float(1.0 / gb))

    def _stage(self, ent, x, y, B):
        V = self.num_visible
        x = torch.as_tensor(x)
        if x.dim() != 2 or tuple(x.shape) != (B, V):
            raise ValueError(f"expected a [{B}, {V}] matrix, got {tuple(x.shape)}")
        y = torch.as_tensor(y)
        if y.dim() == 2:                       # one-hot rows
            y = y.argmax(dim=1)
        if tuple(y.shape) != (B,):
            raise ValueError(f"expected {B} labels, got shape {tuple(y.shape)}")
        ent["x"][:, :V].copy_(x, non_blocking=True)
        ent["y"].copy_(y, non_blocking=True)   # (converts to int32 on the way)

    # ---- Gibbs primitives (:43-60) --------------------------------------------------------------------
    def sample_hidden(self, visible_activations, class_activations):
        L = _lib.load()
        v = self._pad_rows(visible_activations, self.num_visible)
        B = v.shape[0]
        y = self._labels(class_activations, B)
        P = _padded(B, self.num_hidden, self.device)
        self._call(L.qbm_rbm_sample_hidden, self._Wt.data_ptr(), self._U.data_ptr(), self.hidden_bias.data_ptr(), v.data_ptr(),
                   y.data_ptr(), B, self.num_visible, self.num_hidden, self.num_classes, P.data_ptr())
        return P[:, :self.num_hidden]

    def sample_visible(self, hidden_activations):
        L = _lib.load()
        h = self._pad_rows(hidden_activations, self.num_hidden)
        B = h.shape[0]
        P = _padded(B, self.num_visible, self.device)
        self._call(L.qbm_rbm_sample_visible, self._W.data_ptr(), self.visible_bias.data_ptr(), h.data_ptr(), B,
                   self.num_visible, self.num_hidden, P.data_ptr())
        return P[:, :self.num_visible]

    def sample_class(self, hidden_activations):
        L = _lib.load()
        h = self._pad_rows(hidden_activations, self.num_hidden)
        B = h.shape[0]
        P = _padded(B, self.num_classes, self.device)
        self._call(L.qbm_rbm_sample_class, self._U.data_ptr(), self.class_bias.data_ptr(), h.data_ptr(), B, self.num_hidden,
                   self.num_classes, P.data_ptr())
        return P[:, :self.num_classes]

    def sample_class_given_x(self, input_data):
        L = _lib.load()
        x = self._pad_rows(input_data, self.num_visible)
        B = x.shape[0]
        P = _padded(B, self.num_classes, self.device)
        ws = self._workspace(B)
        self._call(L.qbm_rbm_class_given_x, self._Wt.data_ptr(), self._U.data_ptr(), self.hidden_bias.data_ptr(),
                   self.class_bias.data_ptr(), x.data_ptr(), B, self.num_visible, self.num_hidden, self.num_classes,
                   P.data_ptr(), ws.data_ptr(), ws.numel() * 4)
        return P[:, :self.num_classes]

    # ---- training steps -----------------------------------------------------------------------------------
    def _seed64(self):
        return ctypes.c_uint64(int(self.seed) & (2 ** 64 - 1))

    def _disc_launch(self, B, factor, gb, xp, yp, probs_p, pred_p, loss_p, token_host=0, tick_dev_p=None):
        """All launches of one discriminative step on raw device pointers (run eagerly, or once under graph capture).
        Data-parallel: the gradient sums of this shard into the flat buffer, ONE all-reduce, one fused apply."""
        L = _lib.load()
        ws = self._workspace(B)
        V, H, C = self.num_visible, self.num_hidden, self.num_classes
        if self.pg is None:
            self._call(L.qbm_rbm_disc_step, self._W.data_ptr(), self._Wt.data_ptr(), self._U.data_ptr(),
                       self.visible_bias.data_ptr(), self.hidden_bias.data_ptr(), self.class_bias.data_ptr(), xp, yp, B, V, H, C,
                       float(self.learning_rate), float(factor), float(self.sparse_constant), probs_p, pred_p, loss_p,
                       ws.data_ptr(), ws.numel() * 4)
            return
        def grads(gp):
            self._call(L.qbm_rbm_disc_grad, self._Wt.data_ptr(), self._U.data_ptr(), self.hidden_bias.data_ptr(),
                       self.class_bias.data_ptr(), xp, yp, B, V, H, C, gp, probs_p, pred_p, ws.data_ptr(), ws.numel() * 4)

        self._dp_reduce_apply(grads, factor * self.learning_rate / gb, loss_p, gb, token_host, tick_dev_p)

    def _cd1_launch(self, B, gb, xp, yp, step_host, step_dev_p, token_host=0, tick_dev_p=None):
        """All launches of one CD-1 step; draws are keyed by step_host (+ the device counter when step_dev_p is given)."""
        L = _lib.load()
        ws = self._workspace(B)
        V, H, C = self.num_visible, self.num_hidden, self.num_classes
        params = (self._W.data_ptr(), self._Wt.data_ptr(), self._U.data_ptr(), self.visible_bias.data_ptr(),
                  self.hidden_bias.data_ptr(), self.class_bias.data_ptr(), xp, yp, B, V, H, C)
        step = (ctypes.c_uint(step_host),) if step_dev_p is None else (ctypes.c_uint(step_host), step_dev_p)
        if self.pg is None:
            fn = L.qbm_rbm_cd1_step if step_dev_p is None else L.qbm_rbm_cd1_step_dev
            self._call(fn, *params, float(self.learning_rate), float(self.sparse_constant), self._seed64(), *step,
                       ws.data_ptr(), ws.numel() * 4)
            return
        fn = L.qbm_rbm_cd1_grad if step_dev_p is None else L.qbm_rbm_cd1_grad_dev

        def grads(gp):
            self._call(fn, *params, self._seed64(), *step, gp, ws.data_ptr(), ws.numel() * 4)

        self._dp_reduce_apply(grads, self.learning_rate / gb, None, gb, token_host, tick_dev_p)

    def _counters(self):
        """Device counters the captured steps read and advance: [step * world (Philox stream base), step (peer tokens)]."""
        if self._ctr is None:
            self._ctr = torch.zeros(2, dtype=torch.int32, device=self.device)
            self._ctr_inc = torch.tensor([self._world(), 1], dtype=torch.int32, device=self.device)
        return self._ctr

    def _sync_counters(self):
        if self._ctr_val != self._step:               # eager steps in between: resynchronise the device counters
            self._ctr.copy_(torch.tensor([(self._step * self._world()) & 0x3FFFFFFF, self._step & 0x3FFFFFFF], dtype=torch.int32),
                            non_blocking=True)

    def discriminative_training(self, input_data, class_label, factor=1, global_batch=None):
        """:101-146.  Returns (error, predicted, class_probabilities) as CUDA tensors.  With a process group
        ``input_data`` is this rank's shard of a minibatch of ``global_batch`` rows (default: shard x world)."""
        B = len(input_data)
        if B < 2:
            raise ValueError("batch size must be >= 2 (the reference squeezes the batch axis at B = 1, :134)")
        C, lC = self.num_classes, _ld4(self.num_classes)
        gb = float(global_batch if global_batch is not None else B * self._world())
        if self.use_graphs:
            ctr = self._counters()

            def launch(ent):
                self._disc_launch(B, factor, gb, ent["x"].data_ptr(), ent["y"].data_ptr(), ent["out"].data_ptr(),
                                  ent["pred"].data_ptr(), ent["out"].data_ptr() + 4 * B * lC, 1, ctr.data_ptr() + 4)
                ctr.add_(self._ctr_inc)           # part of the graph

            # (data-parallel steps alternate between the two halves of the peer gradient buffer: one capture per parity)
            ent = self._graph_entry(("disc", B, float(self.learning_rate), float(factor), float(self.sparse_constant), gb,
                                     self._step & 1 if self.pg is not None else 0), B, launch)
            if ent is not None:
                self._sync_counters()
                self._stage(ent, input_data, class_label, B)
                ent["graph"].replay()
                self._step += 1
                self._ctr_val = self._step
                out = ent["out"].clone()          # the static buffers are overwritten by the next step
                return out[B * lC], ent["pred"].to(torch.int64), out[:B * lC].view(B, lC)[:, :C]
        x = self._pad_rows(input_data, self.num_visible)
        y = self._labels(class_label, B)
        probs = _padded(B, C, self.device)
        pred = torch.empty(B, dtype=torch.int32, device=self.device)
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        self._disc_launch(B, factor, gb, x.data_ptr(), y.data_ptr(), probs.data_ptr(), pred.data_ptr(), loss.data_ptr(),
                          self._step + 1, None)
        self._step += 1
        return loss[0], pred.to(torch.int64), probs[:, :C]

    def cd1_training(self, input_data, class_label, global_batch=None):
        """One CD-1 step (k = 1) on a minibatch (or this rank's shard of it); parameters updated in place.  Sharded
        minibatches draw from disjoint Philox streams: stream id = step * world + rank."""
        B = len(input_data)
        world, rank = self._world(), self._rank()
        gb = float(global_batch if global_batch is not None else B * world)
        if self.use_graphs:
            ctr = self._counters()

            def launch(ent):
                self._cd1_launch(B, gb, ent["x"].data_ptr(), ent["y"].data_ptr(), rank, ctr.data_ptr(), 1, ctr.data_ptr() + 4)
                ctr.add_(self._ctr_inc)           # part of the graph: the next replay draws from the next streams

            ent = self._graph_entry(("cd1", B, float(self.learning_rate), float(self.sparse_constant), gb,
                                     self._step & 1 if self.pg is not None else 0), B, launch)
            if ent is not None:
                self._sync_counters()
                self._stage(ent, input_data, class_label, B)
                ent["graph"].replay()
                self._step += 1
                self._ctr_val = self._step
                return
        v0 = self._pad_rows(input_data, self.num_visible)
        y0 = self._labels(class_label, B)
        self._cd1_launch(B, gb, v0.data_ptr(), y0.data_ptr(), (self._step * world + rank) & 0x3FFFFFFF, None, self._step + 1, None)
        self._step += 1

    def predict(self, input_data):
        return self.sample_class_given_x(input_data).argmax(dim=1)

    # ---- epoch loop (src/ClassificationRBM.py:159-205) ---------------------------------------------------------------
    def train_rbm(self, train_loader, epochs, cuda=True, validation_loader=None, test_loader=None, method="discriminative",
                  generative_factor=None, discriminative_factor=1):
        """Same loop and return value ``(loss_list, best_validation_model, nll_list)`` as the reference; ``method`` may
        also be ``"cd1"`` (the reference raises NotImplementedError for everything but ``"discriminative"``).  Accuracy on
        ``test_loader`` is appended to ``acc_per_epoch_list`` after every epoch when a loader is given."""
        if method not in ("discriminative", "cd1"):
            raise NotImplementedError(method)
        loss_list, nll_list = [], []
        for _ in range(epochs):
            epoch_error, epoch_nll, nb = 0.0, 0.0, 0
            for batch, labels in train_loader:
                batch = torch.as_tensor(batch).reshape(len(batch), self.num_visible)
                labels = torch.as_tensor(labels).to(self.device)
                if method == "discriminative":
                    err, _, probs = self.discriminative_training(batch, labels, factor=discriminative_factor)
                    nll = -torch.log(probs + 1e-8)[torch.arange(len(batch), device=self.device), labels.long()]
                    epoch_error += float(err)
                    epoch_nll += float(nll.mean())
                else:
                    self.cd1_training(batch, labels)
                nb += 1
            loss_list.append(epoch_error / max(1, nb))
            nll_list.append(epoch_nll / max(1, nb))
            if test_loader is not None:
                correct = total = 0
                for batch, labels in test_loader:
                    batch = torch.as_tensor(batch).reshape(len(batch), self.num_visible)
                    pred = self.predict(batch).cpu()
                    correct += int((pred == torch.as_tensor(labels).long().cpu()).sum())
                    total += len(batch)
                self.acc_per_epoch_list.append(correct / max(1, total))
        return loss_list, self, nll_list
