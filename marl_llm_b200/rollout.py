"""Device-resident replay buffer: host-side mirror of the reference's `ReplayBufferAgent`
(marl_llm/algorithm/utils/buffer_agent.py = BUF) on top of the rollout entry points of the C ABI
(include/swarm_b200.h, group 3).  SURVEY.md §8 f1.

Same constructor, `push`, `sample` and `__len__` as the reference, same ring arithmetic (including the way a push that
does not fit steps the cursor BACK instead of wrapping, BUF:96-99) and the same NumPy-global-RNG calls in `sample`
(BUF:147-157), so that a seeded run draws the same rows.  Differences, all additive:

* arrays live in HBM as fp32 (the reference stores fp64 on the host and casts to fp32 in sample(), BUF:170-173; the
  values handed to the learner are identical);
* `push` also accepts the batched simulator's tensors ([E, dim, n_a], CUDA): E envs are appended env-major in one kernel
  launch; host NumPy inputs of the reference's shapes ([dim, n_a]) are uploaded first;
* `sample(..., to_gpu=True)` returns CUDA tensors without a host round trip (the reference builds them on the host and
  calls .cuda()); `to_gpu=False` returns host tensors like the reference;
* `get_average_rewards` (BUF:179-200) is not mirrored: nothing in the reference calls it and it indexes out of range.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SwarmError, SwarmRolloutBuffers, check


class ReplayBufferAgent:
    def __init__(self, max_steps, num_agents, start_stop_index, state_dim, action_dim, device=0):
        if not torch.cuda.is_available():
            raise SwarmError("the device-resident replay buffer needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.max_steps, self.num_agents = int(max_steps), int(num_agents)
        self.total_length = self.max_steps * self.num_agents                     # BUF:46
        self.state_dim, self.action_dim = int(state_dim), int(action_dim)
        z = lambda d: torch.zeros(self.total_length, d, dtype=torch.float32, device=self.device)   # noqa: E731
        self.obs_buffs, self.ac_buffs, self.ac_prior_buffs = z(self.state_dim), z(self.action_dim), z(self.action_dim)   # BUF:49-51
        self.log_pi_buffs, self.rew_buffs = z(1), z(1)                            # BUF:52-53
        self.next_obs_buffs, self.done_buffs = z(self.state_dim), z(1)            # BUF:54-55
        self.filled_i = 0                                                         # BUF:58
        self.curr_i = 0                                                           # BUF:59
        self.agent_index = start_stop_index                                       # BUF:62
        b = SwarmRolloutBuffers()
        b.struct_size = C.sizeof(SwarmRolloutBuffers)
        b.obs_dim, b.act_dim, b.capacity = self.state_dim, self.action_dim, self.total_length
        b.obs, b.act, b.act_prior = self.obs_buffs.data_ptr(), self.ac_buffs.data_ptr(), self.ac_prior_buffs.data_ptr()
        b.log_pi, b.rew = self.log_pi_buffs.data_ptr(), self.rew_buffs.data_ptr()
        b.next_obs, b.done = self.next_obs_buffs.data_ptr(), self.done_buffs.data_ptr()
        self._b = b
        self.launch_count = 0

    def __len__(self):
        return self.filled_i                                                      # BUF:64-66

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dims, dtype=None):
        """-> contiguous CUDA tensor [E, dims, n_a] (E = 1 for the reference's 2-D host arrays)"""
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if t.dim() == 2:
            t = t.unsqueeze(0)
        assert t.dim() == 3 and t.shape[1] == dims, (tuple(t.shape), dims)
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device).contiguous()

    def push(self, observations_orig, actions_orig, rewards_orig, next_observations_orig, dones_orig, index,
             actions_prior_orig=None, log_pi_orig=None):
        """BUF:67-128.  Arrays are [dim, n_a] (reference) or [E, dim, n_a] (batched simulator); `index` is the agent slice."""
        obs = self._dev(observations_orig, self.state_dim)
        nxt = self._dev(next_observations_orig, self.state_dim)
        E, _, n_a = obs.shape
        out_dt = obs.dtype
        assert out_dt in (torch.float32, torch.float64) and nxt.shape == obs.shape
        nxt = nxt.to(out_dt)
        rew = self._dev(rewards_orig, 1, out_dt)
        done = self._dev(dones_orig, 1, torch.bool).view(torch.uint8)
        act = self._dev(actions_orig, self.action_dim)
        if act.dtype not in (torch.float32, torch.float64):
            act = act.float()
        prior = self._dev(actions_prior_orig, self.action_dim, out_dt) if actions_prior_orig is not None else None
        logpi = self._dev(log_pi_orig, 1, torch.float32) if log_pi_orig is not None else None
        start, stop = index.start or 0, n_a if index.stop is None else index.stop
        data_length = E * len(range(start, stop))                                 # BUF:88-90
        if data_length > self.total_length:
            raise ValueError("one push is larger than the whole buffer")
        if self.curr_i + data_length > self.total_length:                         # BUF:96-99
            rollover = data_length - (self.total_length - self.curr_i)
            self.curr_i -= rollover
        f32 = lambda t: _lib.SWARM_F32 if t.dtype == torch.float32 else _lib.SWARM_F64   # noqa: E731
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None          # noqa: E731
        check(self.lib.swarm_rollout_push(C.byref(self._b), self.curr_i, E, n_a, start, stop, p(obs), p(nxt), p(rew), p(done),
                                          p(prior), f32(obs), p(act), f32(act), p(logpi), self._stream()), "swarm_rollout_push")
        self.launch_count += 1
        self.curr_i += data_length                                                # BUF:116
        if self.filled_i < self.total_length:                                     # BUF:119-120
            self.filled_i += data_length
        if self.curr_i == self.total_length:                                      # BUF:123-124
            self.curr_i = 0

    def sample_indices(self, N):
        """The rows BUF:147-157 draws (NumPy global RNG, same call sequence)."""
        begin_index_range = 3e5                                                   # BUF:147
        begin_index = np.random.randint(0, begin_index_range)
        return np.random.choice(np.arange(begin_index, self.total_length - begin_index_range + begin_index, dtype=np.int32),
                                size=N, replace=False)

    def gather(self, inds, is_prior=False, is_log_pi=False):
        """Rows `inds` (any integer array / tensor) of every array, as CUDA fp32 tensors (BUF:152-165)."""
        # indices that start on the host are validated there (no device round trip); device tensors are trusted to be in
        # range — the kernel clamps rows into the ring instead of reading out of bounds (include/swarm_b200.h)
        if not isinstance(inds, torch.Tensor) or not inds.is_cuda:
            host = np.asarray(inds)
            if host.size and (host.min() < 0 or host.max() >= self.total_length):
                raise IndexError("sample row out of range")
            inds = torch.from_numpy(np.ascontiguousarray(host, dtype=np.int64))
        idx = inds.to(self.device, torch.int64).contiguous()
        n = int(idx.numel())
        o = lambda d: torch.empty(n, d, dtype=torch.float32, device=self.device)   # noqa: E731
        obs, act, rew, nxt, done = o(self.state_dim), o(self.action_dim), o(1), o(self.state_dim), o(1)
        prior = o(self.action_dim) if is_prior else None
        logpi = o(1) if is_log_pi else None
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None          # noqa: E731
        check(self.lib.swarm_rollout_gather(C.byref(self._b), p(idx), n, p(obs), p(act), p(rew), p(nxt), p(done), p(prior),
                                            p(logpi), self._stream()), "swarm_rollout_gather")
        self.launch_count += 1
        return obs, act, rew, nxt, done, prior, logpi

    def sample(self, N, to_gpu=False, is_prior=False, is_log_pi=False):
        """BUF:130-177: 7-tuple (obs, acs, rews, next_obs, dones, prior or None, log_pi or None) of fp32 tensors."""
        out = self.gather(self.sample_indices(N), is_prior, is_log_pi)
        if to_gpu:
            return out
        return tuple(t.cpu() if t is not None else None for t in out)
