"""Time-indexed replay ring for the device-resident rollout loop.

The reference's `ReplayBufferAgent` (marl_llm/algorithm/utils/buffer_agent.py) stores `obs` and `next_obs` of every transition
separately.  Inside one rollout `next_obs` of step t IS `obs` of step t+1, so this ring stores every observation once, as
agent-major rows in a slot per time step:

    obs_ring [T + 1, N, obs_dim]      N = num_envs * n_a rows per slot; slot t = observations before step t, slot T = the last ones
    act / act_prior [T, N, act_dim]   rew / done / log_pi [T, N, 1]

`transition (t, n)`: obs = obs_ring[t, n], next_obs = obs_ring[t + 1, n].  The rows of slot t are written by the POLICY kernel
(`DevicePolicy.step(..., rows_out=ring.slot(t))`: its loader threads hold each agent's observation in registers anyway), so
the 1.5 GB transposes of a conventional push disappear from the loop; only the last slot needs one explicit transpose
(`close()`).  `sample(n)` draws transitions uniformly on the device and returns the reference's 7-tuple
(buffer_agent.py:168-177).  This is an addition next to `marl_llm_b200.rollout.ReplayBufferAgent` (the drop-in mirror), not a
replacement: its sampling rule is uniform over the stored rollout, not the reference's window arithmetic."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SwarmError, SwarmRolloutBuffers, check


class EpisodeRing:
    def __init__(self, steps, num_envs, n_a, obs_dim, act_dim, device=0):
        if not torch.cuda.is_available():
            raise SwarmError("EpisodeRing needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.T, self.E, self.n_a, self.D, self.A = int(steps), int(num_envs), int(n_a), int(obs_dim), int(act_dim)
        self.N = self.E * self.n_a
        z = lambda slots, d: torch.zeros(slots, self.N, d, dtype=torch.float32, device=self.device)   # noqa: E731
        self.obs_ring = z(self.T + 1, self.D)
        self.act, self.act_prior = z(self.T + 1, self.A), z(self.T + 1, self.A)      # one spare slot keeps the ABI's single capacity
        self.rew, self.done, self.log_pi = z(self.T + 1, 1), z(self.T + 1, 1), z(self.T + 1, 1)
        b = SwarmRolloutBuffers()
        b.struct_size = C.sizeof(SwarmRolloutBuffers)
        b.obs_dim, b.act_dim, b.capacity = self.D, self.A, (self.T + 1) * self.N
        b.obs, b.act, b.act_prior = self.obs_ring.data_ptr(), self.act.data_ptr(), self.act_prior.data_ptr()
        b.log_pi, b.rew, b.done, b.next_obs = self.log_pi.data_ptr(), self.rew.data_ptr(), self.done.data_ptr(), None
        self._b = b
        self.filled = 0            # steps recorded (transitions = filled * N once close() has written slot `filled`)
        self.closed = False
        self.launch_count = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def begin(self):
        self.filled, self.closed = 0, False

    def slot(self, t):
        """[N, obs_dim] view of observation slot t: hand it to DevicePolicy.step(rows_out=...)."""
        return self.obs_ring[t]

    def slot_env(self, t):
        """[E, n_a, obs_dim] view of slot t: the observation buffer of an agent-major simulator (`sim.set_obs_buffer`)."""
        return self.obs_ring[t].view(self.E, self.n_a, self.D)

    def begin_direct(self, sim):
        """Start a rollout in which an agent-major simulator writes its observations straight into the slots: its current
        observation is copied into slot 0 once and slot 0 becomes its observation buffer."""
        assert sim.obs_layout == "agent_major" and sim.out_dtype == torch.float32
        self.begin()
        s0 = self.slot_env(0)
        if sim.obs.data_ptr() != s0.data_ptr():
            s0.copy_(sim.obs)
            sim.set_obs_buffer(s0)

    def record(self, t, act, rew, done, act_prior=None, log_pi=None):
        """Small per-agent arrays of step t, from the simulator's layouts ([E, dim, n_a]); one launch, no observation traffic."""
        assert t == self.filled and t < self.T, (t, self.filled, self.T)
        f32 = lambda x: _lib.SWARM_F32 if x.dtype == torch.float32 else _lib.SWARM_F64   # noqa: E731
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None                 # noqa: E731
        assert rew.dtype == torch.float32 and (act_prior is None or act_prior.dtype == torch.float32)
        done8 = done.view(torch.uint8) if done.dtype == torch.bool else done
        check(self.lib.swarm_rollout_push_parts(C.byref(self._b), t * self.N, self.E, self.n_a, 0, self.n_a, None, None, p(rew), p(done8),
                                                p(act_prior), _lib.SWARM_F32, p(act), f32(act), p(log_pi), _lib.SWARM_PUSH_SMALL,
                                                self._stream()), "swarm_rollout_push_parts")
        self.launch_count += 1
        self.filled += 1

    def close(self, last_obs):
        """Transpose the final observation ([E, obs_dim, n_a] fp32) into slot `filled`: next_obs of the last recorded step."""
        assert last_obs.dtype == torch.float32 and last_obs.is_contiguous()
        check(self.lib.swarm_rollout_push_parts(C.byref(self._b), self.filled * self.N, self.E, self.n_a, 0, self.n_a,
                                                C.c_void_p(last_obs.data_ptr()), None, None, None, None, _lib.SWARM_F32, None,
                                                _lib.SWARM_F32, None, _lib.SWARM_PUSH_OBS, self._stream()), "swarm_rollout_push_parts")
        self.launch_count += 1
        self.closed = True

    def __len__(self):
        return self.filled * self.N if self.closed else max(self.filled - 1, 0) * self.N

    def gather(self, rows, is_prior=False, is_log_pi=False):
        """Transitions at flat ring rows (t * N + n) as the reference's 7-tuple of fp32 CUDA tensors."""
        if not isinstance(rows, torch.Tensor) or not rows.is_cuda:      # host indices: validated on the host, no device sync
            host = np.asarray(rows)
            if host.size and (host.min() < 0 or host.max() >= len(self)):
                raise IndexError("transition out of range")
            rows = torch.from_numpy(np.ascontiguousarray(host, dtype=np.int64))
        idx = rows.to(self.device, torch.int64).contiguous()
        n = int(idx.numel())
        o = lambda d: torch.empty(n, d, dtype=torch.float32, device=self.device)   # noqa: E731
        obs, act, rew, nxt, done = o(self.D), o(self.A), o(1), o(self.D), o(1)
        prior, logpi = (o(self.A) if is_prior else None), (o(1) if is_log_pi else None)
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None          # noqa: E731
        check(self.lib.swarm_rollout_gather_ring(C.byref(self._b), p(idx), n, self.N, p(obs), p(act), p(rew), p(nxt), p(done), p(prior),
                                                 p(logpi), self._stream()), "swarm_rollout_gather_ring")
        self.launch_count += 1
        return obs, act, rew, nxt, done, prior, logpi

    def sample(self, n, is_prior=False, is_log_pi=False, generator=None):
        rows = torch.randint(0, len(self), (n,), device=self.device, generator=generator)
        return self.gather(rows, is_prior, is_log_pi)
