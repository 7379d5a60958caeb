"""Duck type of ``neal.SimulatedAnnealingSampler`` (dwave-neal 0.5.9 signature, SURVEY.md Appendix
A.3) whose annealing loop is the sm_100a kernel.  Call sites: src/qubo/sampler.py:22,31-32;
src/model/faster_dqbm.py:102-113,300-301,313.  Stateless, hence picklable (the reference's process
pool pickles ``(bqm, sampler)``, faster_dqbm.py:593)."""
from __future__ import annotations

import numpy as np

from . import dimod_shim as dimod

__version__ = "0.5.9+qbm_b200.shim"


class SimulatedAnnealingSampler:
    parameters = {"beta_range": [], "num_reads": [], "num_sweeps": [], "beta_schedule_type": ["beta_schedule_options"],
                  "seed": [], "interrupt_function": [], "initial_states": [], "initial_states_generator": []}
    properties = {"beta_schedule_options": ("linear", "geometric")}

    def sample(self, bqm, beta_range=None, num_reads=None, num_sweeps=1000, beta_schedule_type="geometric",
               seed=None, interrupt_function=None, initial_states=None, initial_states_generator="random",
               **kwargs):
        from .. import ising, sampler as _sampler
        import torch

        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)} (dwave-neal 0.5.9 signature)")
        if interrupt_function is not None:
            raise NotImplementedError("interrupt_function is not supported by the B200 sampler")
        if not isinstance(num_sweeps, (int, np.integer)) or isinstance(num_sweeps, bool):
            raise TypeError("'num_sweeps' should be a positive integer")
        if num_sweeps < 0:
            raise ValueError("'num_sweeps' should be a positive integer")
        seed = ising.check_seed(seed)
        n = bqm.num_variables
        binary = bqm.vartype is dimod.BINARY
        Q = bqm.to_qubo_matrix() if binary else bqm.change_vartype(dimod.BINARY, inplace=False).to_qubo_matrix()
        qoffset = bqm.offset if binary else bqm.change_vartype(dimod.BINARY, inplace=False).offset

        init = None
        if initial_states is not None:
            arr = dimod._samples_array(initial_states, n)
            init_vt = bqm.vartype
            if isinstance(initial_states, dimod.SampleSet):
                arr, init_vt = initial_states.record.sample, initial_states.vartype
            init01 = (np.asarray(arr) > 0).astype(np.int8) if init_vt is dimod.SPIN else np.asarray(arr).astype(np.int8)
            if num_reads is None:
                num_reads = init01.shape[0]
            if init01.shape[0] < num_reads:
                if initial_states_generator == "none":
                    raise ValueError("insufficient number of initial states given")
                if initial_states_generator == "tile":
                    reps = -(-num_reads // init01.shape[0])
                    init01 = np.tile(init01, (reps, 1))[:num_reads]
                else:  # "random": keep the given ones, fill the rest randomly
                    extra = ising.initial_states_numpy(seed, num_reads - init01.shape[0], n)
                    init01 = np.concatenate([init01, extra], axis=0)
            init = init01[:num_reads]
        if num_reads is None:
            num_reads = 1
        if not isinstance(num_reads, (int, np.integer)) or num_reads < 1:
            raise ValueError("'num_reads' should be a positive integer")
        if n == 0:
            return dimod.SampleSet.from_samples(np.zeros((num_reads, 0), dtype=np.int8), np.zeros(num_reads) + bqm.offset,
                                                bqm.vartype, info={"beta_range": [0.1, 1.0], "beta_schedule_type": beta_schedule_type})
        if seed is None:
            seed = int(np.random.randint(2 ** 31))

        dev = _sampler._require_cuda()
        h, J, _ = ising.qubo_to_ising(Q)
        br = ising.default_beta_range(h, J) if beta_range is None else np.asarray(beta_range, dtype=np.float64).reshape(1, 2)
        betas, spb = ising.beta_schedule(br, int(num_sweeps), beta_schedule_type)
        if init is None:
            init = ising.initial_states_numpy(seed, int(num_reads), n)
        # num_sweeps = 0: neal returns the initial states with their energies (sa_sample short-circuits the empty schedule)
        res = _sampler.sa_sample(torch.from_numpy(J.astype(np.float32)).to(dev), torch.from_numpy(h.astype(np.float32)).to(dev),
                                 torch.from_numpy(betas.astype(np.float32)).to(dev), spb, int(num_reads), seed,
                                 init_states=torch.from_numpy(np.ascontiguousarray(init)[None]).to(dev))
        energies = _sampler.qubo_energies(torch.from_numpy(Q).to(dev), res.states).cpu().numpy()[0] + qoffset
        samples01 = res.states.cpu().numpy()[0]
        info = {"beta_range": [float(br[0, 0]), float(br[0, 1])], "beta_schedule_type": beta_schedule_type}
        if binary:
            return dimod.SampleSet.from_samples(samples01, energies, dimod.BINARY, info=info)
        return dimod.SampleSet.from_samples(2 * samples01.astype(np.int8) - 1, energies, dimod.SPIN, info=info)

    def sample_qubo(self, Q, **parameters):
        """dimod.Sampler mixin: Q is a dict {(u, v): bias} over integer labels 0..n-1."""
        n = 1 + max(max(u, v) for u, v in Q) if Q else 0
        dense = np.zeros((n, n))
        for (u, v), b in Q.items():
            dense[u, v] += b
        return self.sample(dimod.BQM(dense, "BINARY"), **parameters)

    def sample_ising(self, h, J, **parameters):
        """dimod.Sampler mixin: h dict/list of linear biases, J dict {(u, v): coupling}."""
        if not isinstance(h, dict):
            h = dict(enumerate(h))
        return self.sample(dimod.BQM(dict(h), dict(J), 0.0, dimod.SPIN), **parameters)
