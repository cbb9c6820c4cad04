"""Rollout policy on the device: host-side mirror of `DDPGAgent.step` / `MADDPG.step`
(marl_llm/algorithm/utils/agents.py:69-96, marl_llm/algorithm/algorithms/maddpg.py:72-87) on top of group 4 of the C ABI.
SURVEY.md §8 f1: with the simulator on the GPU, evaluating the 192-180-180-180-2 MLP there keeps the rollout loop
(policy -> env.step -> buffer.push) free of host round trips.

`DevicePolicy.step(obs, explore)` takes the simulator's observation tensor ([E, obs_dim, n_a] or the reference's [obs_dim, n_a],
fp32, CUDA) and returns `(actions, log_pi)` in the simulator's action layout ([E, act_dim, n_a] / [act_dim, n_a]) — i.e. the
`.t()` of agents.py:95 is already applied.  Exploration follows agents.py:85-93: one host draw decides between the epsilon
branch (uniform actions for the whole batch) and additive Gaussian noise with clamp; the per-element noise itself comes from a
counter-based generator on the device (seeded, reproducible), not from NumPy's stream."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SwarmError, check


class DevicePolicy:
    def __init__(self, obs_dim, act_dim, hidden_dim=180, device=0, noise_scale=0.1, epsilon=0.0, seed=0, precision="fp32"):
        if not torch.cuda.is_available():
            raise SwarmError("DevicePolicy needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.obs_dim, self.act_dim, self.hidden_dim = int(obs_dim), int(act_dim), int(hidden_dim)
        self.scale, self.epsilon, self.seed = float(noise_scale), float(epsilon), int(seed)   # agents.py:38-44
        self._calls = 0
        h = C.c_void_p()
        check(self.lib.swarm_policy_create(self.device.index, self.obs_dim, self.hidden_dim, self.act_dim, C.byref(h)),
              "swarm_policy_create")
        self._h = h
        self.set_precision(precision)

    def set_precision(self, precision):
        """'fp32': exact path (FFMA, agrees with torch fp32 to rounding).  'f16_tc': persistent tcgen05 kernel, fp16 operands
        with fp32 accumulation in tensor memory (fast mode, ~1e-3 absolute deviation on the tanh output).  'f16x3_tc': the same
        kernel structure with every operand split into fp16 hi + lo (three MMAs per k-step): fp32-accurate (~1e-6)."""
        mode = {"fp32": _lib.SWARM_POLICY_FP32, "f16_tc": _lib.SWARM_POLICY_F16_TC, "f16x3_tc": _lib.SWARM_POLICY_F16X3_TC}[precision]
        check(self.lib.swarm_policy_set_precision(self._h, mode), "swarm_policy_set_precision")
        self.precision = precision

    def close(self):
        if getattr(self, "_h", None):
            self.lib.swarm_policy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd):
        """sd: state dict of the reference's MLPNetwork (fc1..fc4 weight / bias, networks.py:22-25), any device / dtype."""
        arrs = []
        for name, shape in (("fc1", (self.hidden_dim, self.obs_dim)), ("fc2", (self.hidden_dim, self.hidden_dim)),
                            ("fc3", (self.hidden_dim, self.hidden_dim)), ("fc4", (self.act_dim, self.hidden_dim))):
            w = np.ascontiguousarray(sd[name + ".weight"].detach().cpu().numpy(), dtype=np.float32)
            b = np.ascontiguousarray(sd[name + ".bias"].detach().cpu().numpy(), dtype=np.float32)
            assert w.shape == shape and b.shape == (shape[0],), (name, w.shape, shape)
            arrs += [w, b]
        check(self.lib.swarm_policy_load(self._h, *[C.c_void_p(a.ctypes.data) for a in arrs]), "swarm_policy_load")
        return self

    def scale_noise(self, scale):                       # agents.py:60-67
        self.scale = float(scale)

    def step(self, obs, explore=False, out=None, want_log_pi=True, rows_out=None, agent_major=False):
        """agents.py:69-96.  Returns (action, log_pi): [E, act_dim, n_a], [E, 1, n_a] (leading axis dropped for 2-D input).
        agent_major=True: obs is [E, n_a, obs_dim] (a simulator created with obs_layout='agent_major', or a replay-ring slot)."""
        if not (isinstance(obs, torch.Tensor) and obs.is_cuda and obs.dtype == torch.float32):
            raise TypeError("DevicePolicy.step takes the simulator's fp32 CUDA observation tensor")
        squeeze = obs.dim() == 2
        o = (obs.unsqueeze(0) if squeeze else obs).contiguous()
        if agent_major:
            E, n_a, D = o.shape
        else:
            E, D, n_a = o.shape
        assert D == self.obs_dim
        check(self.lib.swarm_policy_obs_layout(self._h, int(bool(agent_major))), "swarm_policy_obs_layout")
        act = out if out is not None else torch.empty(E, self.act_dim, n_a, dtype=torch.float32, device=self.device)
        assert act.is_contiguous() and act.numel() == E * self.act_dim * n_a and act.dtype == torch.float32
        log_pi = torch.empty(E, 1, n_a, dtype=torch.float32, device=self.device) if want_log_pi else None
        mode = 0
        if explore:                                     # agents.py:85-86: ONE draw per step decides the branch for the whole batch
            mode = 2 if np.random.rand() < self.epsilon else 1
        if rows_out is not None:        # agent-major copy of the observations ([E * n_a, obs_dim] fp32): replay storage for free
            assert rows_out.is_cuda and rows_out.dtype == torch.float32 and rows_out.is_contiguous() and rows_out.numel() == E * n_a * D
        check(self.lib.swarm_policy_rows_out(self._h, C.c_void_p(rows_out.data_ptr()) if rows_out is not None else None),
              "swarm_policy_rows_out")
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(self.lib.swarm_policy_step(self._h, C.c_void_p(o.data_ptr()), E, n_a, C.c_void_p(act.data_ptr()),
                                         C.c_void_p(log_pi.data_ptr()) if want_log_pi else None, mode, self.scale, self.seed,
                                         self._calls, stream), "swarm_policy_step")
        self._calls += 1
        if squeeze:
            return act.view(self.act_dim, n_a), (log_pi.view(1, n_a) if want_log_pi else None)
        return act, log_pi

    @property
    def launch_count(self):
        return int(self.lib.swarm_policy_launch_count(self._h))
