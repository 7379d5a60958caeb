"""Sampler boundary B1: a drop-in for ``src/qubo/sampler.py`` whose native
loop is the sm_100a kernel instead of dwave-neal's ``cpu_sa.cpp``.

* :class:`B200SASampler` mirrors ``LocalSASampler`` (src/qubo/sampler.py:19-33): same constructor
  (``num_sweeps``, ``seed``), same ``sample_Q(Q, num_reads) -> float32 [num_reads, n]`` of 0/1 in read
  order, same linear-only shortcut (:13-17, :28-29).  It can be assigned to ``model.sampler``
  (src/model/cdqbm_state.py:55) unchanged.
* :func:`sa_sample` is the batched device-level call under it (torch tensors in, torch tensors out,
  no host synchronisation) used by the training loops and the benchmark.

PyTorch is used only as the owner of device buffers and streams; all arithmetic of the path runs in
libqbm_b200.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, dist as _d, ising


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("qbm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


@dataclass
class SAResult:
    states: torch.Tensor            # int8 [batch_q, num_reads, n] of 0/1, read order
    accepted: torch.Tensor | None   # uint64-as-int64 [2]: accepted flips, proposals (when counted)


def sa_sample(J: torch.Tensor, h: torch.Tensor, betas: torch.Tensor, sweeps_per_beta: int, num_reads: int,
              seed: int, chain_offset: int = 0, init_states: torch.Tensor | None = None,
              count: bool = False, flags: int = 0, workspace: torch.Tensor | None = None,
              out: torch.Tensor | None = None) -> SAResult:
    """Anneal ``num_reads`` chains for each of ``batch_q`` spin models on the current CUDA stream.

    ``J`` float32 [batch_q, n, n] (symmetric, zero diagonal), ``h`` float32 [batch_q, n],
    ``betas`` float32 [batch_q, num_betas] or [1, num_betas] / [num_betas] (shared schedule).
    ``flags`` is passed to ``qbm_sa_sample`` (include/qbm_b200.h): 0 = the library's choice -- the warp-per-chain kernel,
    for n > 896 preceded by the chain-tile kernel over the hot sweeps (two-phase schedule); 64 = never two-phase,
    16 / 32 = the chain-tile / chains-per-warp kernel for the whole schedule.  Every choice returns the same states.
    """
    L = _lib.load()
    if J.dim() == 2:
        J = J[None]
    if h.dim() == 1:
        h = h[None]
    if betas.dim() == 1:
        betas = betas[None]
    if not (J.is_cuda and h.is_cuda and betas.is_cuda):
        raise ValueError("sa_sample: J, h and betas must be CUDA tensors")
    if J.dtype != torch.float32 or h.dtype != torch.float32 or betas.dtype != torch.float32:
        raise ValueError("sa_sample: J, h and betas must be float32")
    J = J.contiguous(); h = h.contiguous(); betas = betas.contiguous()
    bq, n, n2 = J.shape
    if n != n2 or h.shape != (bq, n):
        raise ValueError(f"sa_sample: inconsistent shapes J={tuple(J.shape)} h={tuple(h.shape)}")
    if betas.shape[0] not in (1, bq):
        raise ValueError("sa_sample: betas must have 1 or batch_q rows")
    num_betas = betas.shape[1]
    beta_stride = 0 if betas.shape[0] == 1 else num_betas
    dev = J.device
    if num_betas == 0:
        # neal with num_sweeps = 0 returns the initial states: nothing to anneal, so nothing to launch
        if init_states is None:
            raise ValueError("sa_sample: an empty schedule needs explicit init_states (there is nothing to anneal)")
        if init_states.shape != (bq, num_reads, n) or init_states.dtype != torch.int8:
            raise ValueError("sa_sample: init_states must be int8 [batch_q, num_reads, n]")
        states = init_states.contiguous().clone() if out is None else out.copy_(init_states)
        return SAResult(states=states, accepted=torch.zeros(2, dtype=torch.int64, device=dev) if count else None)
    # the larger workspace lets the library run its two-phase schedule where that is faster (n > 896)
    need = L.qbm_sa_workspace_bytes_two_phase(n, bq, int(num_reads))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 15) // 16 * 4, dtype=torch.float32, device=dev)
    if out is None:
        out = torch.empty((bq, num_reads, n), dtype=torch.int8, device=dev)
    elif out.shape != (bq, num_reads, n) or out.dtype != torch.int8 or not out.is_contiguous():
        raise ValueError("sa_sample: `out` must be a contiguous int8 [batch_q, num_reads, n] tensor")
    init_ptr = None
    if init_states is not None:
        if init_states.shape != (bq, num_reads, n) or init_states.dtype != torch.int8:
            raise ValueError("sa_sample: init_states must be int8 [batch_q, num_reads, n]")
        init_states = init_states.contiguous()
        init_ptr = init_states.data_ptr()
    counters = torch.zeros(2, dtype=torch.int64, device=dev) if count else None
    with torch.cuda.device(dev):
        rc = L.qbm_sa_sample(J.data_ptr(), h.data_ptr(), n, n, bq, betas.data_ptr(), beta_stride, num_betas,
                             int(sweeps_per_beta), int(num_reads), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                             ctypes.c_uint64(int(chain_offset)), init_ptr, out.data_ptr(),
                             counters.data_ptr() if counters is not None else None,
                             workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                             int(flags), _stream_ptr(dev))
    _lib.check(rc)
    return SAResult(states=out, accepted=counters)


def qubo_energies(Q: torch.Tensor, states: torch.Tensor) -> torch.Tensor:
    """float64 [batch_q, R] energies x^T Q x of int8 0/1 ``states`` [batch_q, R, n] (K2)."""
    L = _lib.load()
    if Q.dim() == 2:
        Q = Q[None]
    if states.dim() == 2:
        states = states[None]
    if Q.dtype != torch.float64 or states.dtype != torch.int8 or not Q.is_cuda or not states.is_cuda:
        raise ValueError("qubo_energies: Q must be CUDA float64 and states CUDA int8")
    Q = Q.contiguous(); states = states.contiguous()
    bq, n, _ = Q.shape
    if states.shape[0] != bq or states.shape[2] != n:
        raise ValueError(f"qubo_energies: inconsistent shapes Q={tuple(Q.shape)} states={tuple(states.shape)}")
    R = states.shape[1]
    out = torch.empty((bq, R), dtype=torch.float64, device=Q.device)
    with torch.cuda.device(Q.device):
        rc = L.qbm_qubo_energy(Q.data_ptr(), n, bq, states.data_ptr(), R, out.data_ptr(), _stream_ptr(Q.device))
    _lib.check(rc)
    return out


def phase_stats(states: torch.Tensor, second: bool = True):
    """(mean float64 [batch_q, n], second float64 [batch_q, n, n] or None) of int8 0/1 states (K3)."""
    L = _lib.load()
    if states.dim() == 2:
        states = states[None]
    if states.dtype != torch.int8 or not states.is_cuda:
        raise ValueError("phase_stats: states must be a CUDA int8 tensor")
    states = states.contiguous()
    bq, R, n = states.shape
    dev = states.device
    mean = torch.empty((bq, n), dtype=torch.float64, device=dev)
    sec = torch.empty((bq, n, n), dtype=torch.float64, device=dev) if second else None
    need = L.qbm_phase_stats_workspace_bytes(bq, R, n)
    ws = torch.empty((need + 3) // 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.qbm_phase_stats(states.data_ptr(), bq, R, n, mean.data_ptr(), sec.data_ptr() if second else None,
                               ws.data_ptr(), ws.numel() * 4, _stream_ptr(dev))
    _lib.check(rc)
    return mean, sec


def qubo_to_ising_device(Q: torch.Tensor):
    """K0 on the device: (J f32 [B,n,n], h f32 [B,n], offset f64 [B], range f64 [B,2])."""
    L = _lib.load()
    if Q.dim() == 2:
        Q = Q[None]
    if Q.dtype != torch.float64 or not Q.is_cuda:
        raise ValueError("qubo_to_ising_device: Q must be a CUDA float64 tensor")
    Q = Q.contiguous()
    B, n, _ = Q.shape
    if n > 4096:
        raise ValueError(f"qubo_to_ising_device: n={n} exceeds the 4096 variables the row-sum kernel is built for")
    dev = Q.device
    J = torch.empty((B, n, n), dtype=torch.float32, device=dev)
    h = torch.empty((B, n), dtype=torch.float32, device=dev)
    off = torch.empty(B, dtype=torch.float64, device=dev)
    rng = torch.empty((B, 2), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = L.qbm_qubo_to_ising(Q.data_ptr(), n, B, J.data_ptr(), h.data_ptr(), off.data_ptr(), rng.data_ptr(),
                                 _stream_ptr(dev))
    _lib.check(rc)
    return J, h, off, rng


# ------------------------------------------------------------------------------------------------
# host-level call: numpy QUBOs in, numpy samples out (what the reference's call sites exchange)
# ------------------------------------------------------------------------------------------------
# host <-> device copies of the host-buffer API go through page-locked memory: a pageable copy is a synchronous bounce
# copy of the driver at ~2 GB/s on these hosts (measured: 95 ms for the 205 MB of samples of the 1e5-read job), a
# pinned one a DMA at link speed.  torch's caching host allocator keeps the pinned blocks across calls.
_PIN_MIN_BYTES = 1 << 20


def _to_device(a: np.ndarray, dev) -> torch.Tensor:
    t = torch.from_numpy(a)
    if a.nbytes < _PIN_MIN_BYTES:
        return t.to(dev)
    pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    pin.copy_(t)
    out = pin.to(dev, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()       # the staging block may be reused as soon as it is released
    return out


def _to_host(t: torch.Tensor) -> np.ndarray:
    """numpy copy of a device tensor; large ones land in (and stay backed by) page-locked memory"""
    if t.numel() * t.element_size() < _PIN_MIN_BYTES:
        return t.cpu().numpy()
    pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    pin.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return pin.numpy()


def _all_gather_reads(local: torch.Tensor, num_reads: int, group) -> torch.Tensor:
    """Concatenate the read shards of all ranks along dim 1 (reads): [B, hi - lo, ...] -> [B, num_reads, ...]."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = max(b - a for a, b in (_d.shard_range(num_reads, world, r) for r in range(world)))
    pad = torch.zeros((local.shape[0], width) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:, :b - a] for r, (a, b) in ((r, _d.shard_range(num_reads, world, r)) for r in range(world))], dim=1)


def sample_qubo_batch(Q: np.ndarray, num_reads: int, num_sweeps: int = 1000, seed=None, beta_range=None,
                      beta_schedule_type: str = "geometric", initial_states_generator: str = "numpy",
                      device=None, return_energy: bool = True, chain_offset: int = 0, process_group=None,
                      src_rank: int | None = None):
    """Sample a batch of dense QUBOs ``[B, n, n]`` (float64, host).  Host logic (spin conversion, beta
    range, schedule, initial states) follows neal/dimod in float64 (:mod:`ising`); the annealing, the
    energies and nothing else run on the GPU.

    With ``process_group`` (one process per GPU) the reads are sharded over the ranks -- rank r anneals the
    contiguous block ``dist.shard_range(num_reads, world, r)`` keyed by the global read index, so the result does not
    depend on the number of GPUs -- and all-gathered, so that every rank returns all ``num_reads`` samples in read order.
    A ``seed`` of None is drawn on rank 0 and broadcast, so that the shards belong to one reproducible call.  With
    ``src_rank`` the problem is taken from that rank only: its ``Q`` is broadcast over the group on the device (NCCL; 33.6 MB
    of float64 at n = 2048) and the ``Q`` the other ranks pass is used for its shape alone.

    Returns ``(samples int8 [B, R, n] numpy, energies float64 [B, R] numpy or None, info dict)``.
    """
    dev = _require_cuda(device)
    Q = np.asarray(Q, dtype=np.float64)
    single = Q.ndim == 2
    if single:
        Q = Q[None]
    B, n, _ = Q.shape
    seed = ising.check_seed(seed)
    if seed is None:
        seed = int(np.random.randint(2 ** 31))
        if process_group is not None:
            import torch.distributed as dist
            box = [seed]
            dist.broadcast_object_list(box, src=dist.get_global_rank(process_group, 0), group=process_group)
            seed = int(box[0])
    # BINARY -> SPIN and the two reductions of neal's beta rule run on the device (K0 reproduces the float64 host
    # formulas of ising.py bit for bit, numpy's summation order included); the schedule itself is built on the host from
    # those two numbers exactly as neal does (np.geomspace)
    Qd = _to_device(Q, dev)
    if process_group is not None and src_rank is not None:
        import torch.distributed as dist
        dist.broadcast(Qd, src=dist.get_global_rank(process_group, int(src_rank)), group=process_group)
    Jd, hd, off_d, rng_d = qubo_to_ising_device(Qd)
    host3 = torch.cat([off_d[:, None], rng_d], dim=1).cpu().numpy()       # one device->host copy for offset + both reductions
    offset = host3[:, 0].copy()
    if beta_range is None:
        r = host3[:, 1:]
        br = ising.beta_range_from_reductions(r[:, 0], r[:, 1])
    else:
        br = np.broadcast_to(np.asarray(beta_range, dtype=np.float64), (B, 2))
    betas, spb = ising.beta_schedule(br, num_sweeps, beta_schedule_type)
    bd = torch.from_numpy(betas.astype(np.float32)).to(dev, non_blocking=True)
    lo, hi = 0, int(num_reads)
    if process_group is not None:
        import torch.distributed as dist
        lo, hi = _d.shard_range(int(num_reads), dist.get_world_size(process_group), dist.get_rank(process_group))
        if B > 1:
            raise ValueError("read sharding over a process group supports one problem per call")
    init = None
    if initial_states_generator == "numpy":
        # the reference passes the same seed on every call, so every problem of the batch starts
        # from the same RandomState(seed) draw
        st0 = ising.initial_states_numpy(seed, num_reads, n)[lo:hi]
        init = torch.from_numpy(np.broadcast_to(st0, (B, hi - lo, n)).copy()).to(dev)
    elif initial_states_generator != "philox":
        raise ValueError("initial_states_generator must be 'numpy' or 'philox'")
    states = torch.empty((B, 0, n), dtype=torch.int8, device=dev)
    if hi > lo:
        states = sa_sample(Jd, hd, bd, spb, hi - lo, seed, chain_offset=chain_offset + lo, init_states=init).states
    energies = None
    if return_energy:
        energies = qubo_energies(Qd, states) if hi > lo else torch.empty((B, 0), dtype=torch.float64, device=dev)
    if process_group is not None:
        states = _all_gather_reads(states, int(num_reads), process_group)
        if energies is not None:
            energies = _all_gather_reads(energies, int(num_reads), process_group)
    if energies is not None:
        energies = _to_host(energies)
    samples = _to_host(states)
    info = {"beta_range": br.tolist() if not single else br[0].tolist(), "beta_schedule_type": beta_schedule_type,
            "num_betas": int(betas.shape[1]), "sweeps_per_beta": spb, "offset": offset}
    return samples, energies, info


class B200SASampler:
    """Drop-in for ``LocalSASampler`` (src/qubo/sampler.py:19-33)."""

    def __init__(self, num_sweeps: int = 1000, seed: int | None = None, initial_states_generator: str = "numpy",
                 device=None, process_group=None):
        self.num_sweeps = int(num_sweeps)
        self.seed = seed
        self.initial_states_generator = initial_states_generator
        self.device = device
        self.process_group = process_group      # reads sharded over the ranks, all samples returned on every rank

    def sample_Q(self, Q: np.ndarray, num_reads: int) -> np.ndarray:
        Q = np.asarray(Q, dtype=np.float64)
        if Q.ndim != 2 or Q.shape[0] != Q.shape[1]:
            raise ValueError(f"Q must be a square matrix, got shape {Q.shape}")
        if bool(ising.is_linear_only(Q)[0]):
            return _solve_linear_only(Q, int(num_reads), self.seed)
        samples, _, _ = sample_qubo_batch(Q, int(num_reads), self.num_sweeps, self.seed,
                                          initial_states_generator=self.initial_states_generator,
                                          device=self.device, return_energy=False, process_group=self.process_group)
        return _as_float32(samples[0])


def _as_float32(states: np.ndarray) -> np.ndarray:
    """int8 samples -> the float32 matrix the reference's sampler returns (src/qubo/sampler.py:33).  At the 1e5-read job this
    is 2e8 values and 819 MB of fresh pages: one core needs ~0.25 s for it (a quarter of the whole call on eight GPUs; measured
    alternatives -- converting on the device and copying four times the bytes, torch's CPU conversion under torchrun's
    OMP_NUM_THREADS=1 -- are slower), so large results are converted in row blocks on a few threads (numpy releases the GIL)."""
    if states.size < (1 << 24):
        return states.astype(np.float32)
    import os
    from concurrent.futures import ThreadPoolExecutor
    ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    threads = max(1, min(8, (os.cpu_count() or 1) // max(1, ranks_here)))
    out = np.empty(states.shape, dtype=np.float32)
    rows = states.shape[0]
    step = (rows + threads - 1) // threads

    def work(i):
        out[i:i + step] = states[i:i + step]

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(0, rows, step)))
    return out


def _solve_linear_only(Q: np.ndarray, num_reads: int, seed) -> np.ndarray:
    """src/qubo/sampler.py:13-17: the argmin solution replicated ``num_reads`` times (host-side branch;
    not a thermal sample -- SURVEY.md Appendix B Q7)."""
    rng = np.random.default_rng(seed)
    n = Q.shape[0]
    sol = np.zeros(n, dtype=np.float32)
    for v in range(n):
        hv = Q[v, v]
        sol[v] = 1 if hv < 0 else (0 if hv > 0 else int(rng.integers(0, 2)))
    return np.tile(sol, (num_reads, 1)).astype(np.float32)
