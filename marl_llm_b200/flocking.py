"""FlockingSwarm variant on the shared pair core (VARIANTS.md §3).

The reference registers `FlockingSwarm-v0` (cus_gym/gym/envs/__init__.py:7-12) but ships no source for it, so this is a
SPECIFIED variant: no oracle exists and parity is unpinned.  Everything it shares with the assembly env — pair geometry, ball-ball
spring, walls / periodic wrap, integrator, k nearest neighbours, observation head — is literally the assembly step's first-half
kernel (`k_step<PH=1>`), which the assembly parity suite pins to the reference; only the Reynolds reward (`k_flock_reward`) is new.

    sim = BatchedFlockingSim(num_envs, n_a, r_avoid=0.26)
    obs = sim.reset(seed=0)                      # [E, 4 * (6 + self), n_a]
    obs, rew, done, info = sim.step(act)         # act [E, 2, n_a] CUDA tensor (fp32 / fp64)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SwarmBuffers, SwarmConfig, SwarmError, check

TOPO_NEI_MAX = 6
W_COLLISION, W_ALIGN, W_SPACING = 1.0, 0.5, 0.5          # reward weights of VARIANTS.md §3 (k_flock_reward launches with the same)


class BatchedFlockingSim:
    def __init__(self, num_envs, n_a, r_avoid=0.26, *, device=0, out_dtype=torch.float32, is_con_self_state=True, is_periodic=False,
                 d_sen=0.4, size_a=0.035, k_ball=30.0, k_wall=100.0, c_wall=5.0, dt=0.1, vel_max=0.8, mass=1.0, half_width=2.4,
                 half_height=2.4):
        if not torch.cuda.is_available():
            raise SwarmError("BatchedFlockingSim needs a CUDA device; there is no CPU fallback")
        if n_a > 128:
            raise ValueError("the flocking variant supports n_a <= 128")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.E, self.n_a, self.out_dtype = int(num_envs), int(n_a), out_dtype
        self.r_avoid, self.d_sen, self.half = float(r_avoid), float(d_sen), (float(half_width), float(half_height))
        cfg = SwarmConfig()
        cfg.struct_size = C.sizeof(SwarmConfig)
        cfg.device, cfg.num_envs, cfg.n_a, cfg.n_g_max = self.device.index, self.E, self.n_a, 32
        cfg.topo_nei_max, cfg.num_obs_grid_max, cfg.num_occupied_grid_max = TOPO_NEI_MAX, 80, 200
        cfg.is_con_self_state, cfg.is_periodic, cfg.want_prior = int(is_con_self_state), int(is_periodic), 0
        cfg.out_dtype = _lib.SWARM_F32 if out_dtype == torch.float32 else _lib.SWARM_F64
        cfg.variant = _lib.SWARM_VARIANT_FLOCKING
        cfg.d_sen, cfg.r_avoid, cfg.size_a = d_sen, r_avoid, size_a
        cfg.k_ball, cfg.k_wall, cfg.c_wall, cfg.dt, cfg.vel_max, cfg.mass = k_ball, k_wall, c_wall, dt, vel_max, mass
        cfg.boundary_pos[:] = [-half_width, half_height, half_width, -half_height]
        self.obs_dim = int(self.lib.swarm_obs_dim(C.byref(cfg)))
        E, n, dev = self.E, self.n_a, self.device
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)   # noqa: E731
        self.p, self.dp = z(E, 2, n, dtype=torch.float64), z(E, 2, n, dtype=torch.float64)
        self.obs, self.reward = z(E, self.obs_dim, n, dtype=out_dtype), z(E, 1, n, dtype=out_dtype)
        self.done = z(E, 1, n, dtype=torch.bool)
        self.neighbor_index = z(E, n, TOPO_NEI_MAX, dtype=torch.int32).fill_(-1)
        self.in_flags = z(E, n, dtype=torch.int32)
        # the assembly handle's grid-side buffers are unused by this variant: tiny dummies
        self._dummy = [z(E, 32, 2, dtype=torch.float64), z(E, dtype=torch.int32), z(E, 1, 4, dtype=torch.float32), z(E, 2, dtype=torch.float64),
                       z(E, dtype=torch.float64), z(E, 2, n, dtype=out_dtype), z(E, 2, n, dtype=out_dtype), z(E, n, dtype=torch.int32)]
        buf = SwarmBuffers()
        buf.struct_size = C.sizeof(SwarmBuffers)
        buf.p, buf.dp, buf.obs, buf.reward = self.p.data_ptr(), self.dp.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr()
        buf.neighbor_index, buf.in_flags = self.neighbor_index.data_ptr(), self.in_flags.data_ptr()
        d = self._dummy
        buf.grid, buf.n_g, buf.word_box, buf.frame, buf.in_thresh = (t.data_ptr() for t in d[:5])
        buf.a_prior[0], buf.a_prior[1], buf.nearest_cell = d[5].data_ptr(), d[6].data_ptr(), d[7].data_ptr()
        h = C.c_void_p()
        check(self.lib.swarm_create(C.byref(cfg), C.byref(buf), C.byref(h)), "swarm_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.swarm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_state(self, p, dp):
        self.p.copy_(torch.as_tensor(p, dtype=torch.float64).reshape(self.E, 2, self.n_a))
        self.dp.copy_(torch.as_tensor(dp, dtype=torch.float64).reshape(self.E, 2, self.n_a))

    def observe(self):
        check(self.lib.swarm_flock_observe(self._h, self._stream()), "swarm_flock_observe")
        return self.obs

    def reset(self, seed=0):
        """Uniform positions in the arena, velocities U(-0.5, 0.5) (the assembly env's wide spawn, ENV:203-205, 215)."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        hw, hh = self.half
        u = torch.rand(self.E, 2, self.n_a, dtype=torch.float64, device=self.device, generator=g)
        self.p[:, 0] = -hw + 2 * hw * u[:, 0]; self.p[:, 1] = -hh + 2 * hh * u[:, 1]
        self.dp.copy_(torch.rand(self.E, 2, self.n_a, dtype=torch.float64, device=self.device, generator=g) - 0.5)
        return self.observe()

    def step(self, act):
        if not (isinstance(act, torch.Tensor) and act.is_cuda):
            raise TypeError("step() takes a CUDA tensor")
        if act.dtype not in (torch.float32, torch.float64):
            act = act.to(torch.float32)
        act = act.contiguous()
        assert act.numel() == self.E * 2 * self.n_a
        dt = _lib.SWARM_F32 if act.dtype == torch.float32 else _lib.SWARM_F64
        check(self.lib.swarm_flock_step(self._h, C.c_void_p(act.data_ptr()), dt, self._stream()), "swarm_flock_step")
        return self.obs, self.reward, self.done, None


def reward_reference(p, dp, nbr, r_avoid, periodic=False, half=(2.4, 2.4)):
    """NumPy evaluation of the reward of VARIANTS.md §3 for ONE env (the specification the kernel is tested against; there is no
    reference implementation of this variant)."""
    n_a = p.shape[1]
    out = np.zeros(n_a)
    for i in range(n_a):
        js = [j for j in nbr[i] if j >= 0]
        if not js:
            continue
        rel = p[:, js] - p[:, [i]]
        if periodic:
            for k, h in enumerate(half):
                rel[k] = np.where(rel[k] < -h, rel[k] + 2 * h, np.where(rel[k] > h, rel[k] - 2 * h, rel[k]))
        d = np.sqrt(rel[0] * rel[0] + rel[1] * rel[1])
        align = np.linalg.norm(dp[:, js].mean(axis=1) - dp[:, i])
        out[i] = -W_COLLISION * float((d < r_avoid).any()) - W_ALIGN * align - W_SPACING * abs(d.mean() - 2 * r_avoid)
    return out
