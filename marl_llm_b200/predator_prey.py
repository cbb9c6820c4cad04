"""PredatorPreySwarm variant on the pair core (VARIANTS.md §4).

The reference registers `PredatorPreySwarm-v0` (cus_gym/gym/envs/__init__.py:14-19) but ships no source for it, so this is a
SPECIFIED variant: no oracle exists and parity is unpinned.  Dynamics are the assembly step's (spring ENV:442-457 + CPP:775-807,
walls CPP:835-846 + ENV:517-518, integrator ENV:631-652) with a per-type velocity clip; the observation is the assembly head
(CPP:102-126) over the 6 nearest agents of the own type and of the other type.  `step_reference` below is the NumPy statement of
that specification for ONE env — the thing the kernel (`k_pp_step`) is tested against.

    sim = BatchedPredatorPreySim(num_envs, n_p=3, n_e=10)
    obs = sim.reset(seed=0)                      # [E, 4 * (12 + self), n]
    obs, rew, done, info = sim.step(act)         # act [E, 2, n] CUDA tensor (fp32 / fp64); scripted types ignore their rows
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SwarmError, SwarmPPBuffers, SwarmPPConfig, check

TOPO_NEI_MAX = 6
STRATEGIES = {"input": _lib.SWARM_PP_INPUT, "static": _lib.SWARM_PP_STATIC, "random": _lib.SWARM_PP_RANDOM, "nearest": _lib.SWARM_PP_NEAREST}


class BatchedPredatorPreySim:
    def __init__(self, num_envs, n_p, n_e, *, device=0, out_dtype=torch.float32, is_con_self_state=True, is_periodic=False,
                 billiards=False, pursuer_strategy="input", escaper_strategy="input", d_sen=0.4, size_a=0.035, k_ball=30.0,
                 k_wall=100.0, c_wall=5.0, dt=0.1, vel_max_p=0.8, vel_max_e=1.0, mass=1.0, half_width=2.4, half_height=2.4, seed=0):
        if not torch.cuda.is_available():
            raise SwarmError("BatchedPredatorPreySim needs a CUDA device; there is no CPU fallback")
        if not 1 <= n_p + n_e <= 128:
            raise ValueError("the predator-prey variant supports 1 <= n_p + n_e <= 128")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.E, self.n_p, self.n_e, self.n = int(num_envs), int(n_p), int(n_e), int(n_p + n_e)
        self.half = (float(half_width), float(half_height))
        cfg = SwarmPPConfig()
        cfg.struct_size = C.sizeof(SwarmPPConfig)
        cfg.device, cfg.num_envs, cfg.n_p, cfg.n_e = self.device.index, self.E, self.n_p, self.n_e
        cfg.is_con_self_state, cfg.is_periodic, cfg.billiards = int(is_con_self_state), int(is_periodic), int(billiards)
        cfg.out_dtype = _lib.SWARM_F32 if out_dtype == torch.float32 else _lib.SWARM_F64
        cfg.strategy_p, cfg.strategy_e = STRATEGIES[pursuer_strategy], STRATEGIES[escaper_strategy]
        cfg.d_sen, cfg.size_a, cfg.k_ball, cfg.k_wall, cfg.c_wall = d_sen, size_a, k_ball, k_wall, c_wall
        cfg.dt, cfg.vel_max_p, cfg.vel_max_e, cfg.mass, cfg.seed = dt, vel_max_p, vel_max_e, mass, int(seed)
        cfg.boundary_pos[:] = [-half_width, half_height, half_width, -half_height]
        self.cfg = cfg
        self.obs_dim = int(self.lib.swarm_pp_obs_dim(C.byref(cfg)))
        E, n, dev = self.E, self.n, self.device
        self.p = torch.zeros(E, 2, n, dtype=torch.float64, device=dev)
        self.dp = torch.zeros(E, 2, n, dtype=torch.float64, device=dev)
        self.obs = torch.zeros(E, self.obs_dim, n, dtype=out_dtype, device=dev)
        self.reward = torch.zeros(E, 1, n, dtype=out_dtype, device=dev)
        self.done = torch.zeros(E, 1, n, dtype=torch.bool, device=dev)
        self.neighbor_index = torch.full((E, n, 2 * TOPO_NEI_MAX), -1, dtype=torch.int32, device=dev)
        buf = SwarmPPBuffers()
        buf.struct_size = C.sizeof(SwarmPPBuffers)
        buf.p, buf.dp, buf.obs, buf.reward = self.p.data_ptr(), self.dp.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr()
        buf.neighbor_index = self.neighbor_index.data_ptr()
        self.buf = buf
        self.step_index = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_state(self, p, dp):
        self.p.copy_(torch.as_tensor(p, dtype=torch.float64).reshape(self.E, 2, self.n))
        self.dp.copy_(torch.as_tensor(dp, dtype=torch.float64).reshape(self.E, 2, self.n))

    def observe(self):
        check(self.lib.swarm_pp_observe(C.byref(self.cfg), C.byref(self.buf), self._stream()), "swarm_pp_observe")
        return self.obs

    def reset(self, seed=0):
        """Uniform positions in the arena, velocities U(-0.5, 0.5) (the assembly env's wide spawn, ENV:203-205, 215)."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        hw, hh = self.half
        u = torch.rand(self.E, 2, self.n, dtype=torch.float64, device=self.device, generator=g)
        self.p[:, 0] = -hw + 2 * hw * u[:, 0]; self.p[:, 1] = -hh + 2 * hh * u[:, 1]
        self.dp.copy_(torch.rand(self.E, 2, self.n, dtype=torch.float64, device=self.device, generator=g) - 0.5)
        self.step_index = 0
        return self.observe()

    def step(self, act=None):
        need = _lib.SWARM_PP_INPUT in (self.cfg.strategy_p, self.cfg.strategy_e)
        ptr, dt = None, _lib.SWARM_F32
        if need:
            if not (isinstance(act, torch.Tensor) and act.is_cuda):
                raise TypeError("step() takes a CUDA tensor when a side is driven by input actions")
            if act.dtype not in (torch.float32, torch.float64):
                act = act.to(torch.float32)
            act = act.contiguous()
            assert act.numel() == self.E * 2 * self.n
            ptr, dt = C.c_void_p(act.data_ptr()), (_lib.SWARM_F32 if act.dtype == torch.float32 else _lib.SWARM_F64)
        check(self.lib.swarm_pp_step(C.byref(self.cfg), C.byref(self.buf), ptr, dt, self.step_index, self._stream()), "swarm_pp_step")
        self.step_index += 1
        return self.obs, self.reward, self.done, None


def _wrap(rel, half):
    for k, h in enumerate(half):
        rel[k] = np.where(rel[k] < -h, rel[k] + 2 * h, np.where(rel[k] > h, rel[k] - 2 * h, rel[k]))
    return rel


def step_reference(p, dp, act, n_p, *, dyn=True, periodic=False, billiards=False, self_state=True, d_sen=0.4, size_a=0.035,
                   k_ball=30.0, k_wall=100.0, c_wall=5.0, dt=0.1, vel_max_p=0.8, vel_max_e=1.0, mass=1.0, half=(2.4, 2.4)):
    """NumPy statement of VARIANTS.md §4 for ONE env with input actions: returns (p, dp, obs, reward, neighbor_index).  Plain
    float64 loops in the order the specification gives (partners ascending, sums left to right) so that the kernel can be compared
    bit for bit."""
    p, dp = np.array(p, dtype=np.float64), np.array(dp, dtype=np.float64)
    n = p.shape[1]
    hw, hh = half
    bx_min, by_max, bx_max, by_min = -hw, hh, hw, -hh
    is_p = np.arange(n) < n_p
    if dyn:
        a = np.asarray(act, dtype=np.float64)
        new_p, new_dp = p.copy(), dp.copy()
        for i in range(n):
            x, y, vx, vy = p[0, i], p[1, i], dp[0, i], dp[1, i]
            sfx = sfy = 0.0
            for k in range(n):
                if k == i:
                    continue
                dx, dy = p[0, k] - x, p[1, k] - y
                d = np.sqrt(dx * dx + dy * dy)
                if d - 2 * size_a < 0:
                    aa = abs(d - (size_a + size_a)) * k_ball
                    sfx = sfx + aa * ((x - p[0, k]) / d)
                    sfy = sfy + aa * ((y - p[1, k]) / d)
            g = [x - size_a - bx_min, by_max - (y + size_a), bx_max - (x + size_a), y - size_a - by_min]
            m = [abs(v) if v < 0 else 0.0 for v in g]
            sfwx, sfwy = (m[0] - m[2]) * k_wall, (-m[1] + m[3]) * k_wall
            w = [vx if g[0] < 0 else 0.0, vy if g[1] < 0 else 0.0, vx if g[2] < 0 else 0.0, vy if g[3] < 0 else 0.0]
            dfwx, dfwy = (-w[0] - w[2]) * c_wall, (-w[1] - w[3]) * c_wall
            if periodic or billiards:
                fx, fy = a[0, i] + sfx, a[1, i] + sfy
            else:
                fx, fy = ((a[0, i] + sfx) + sfwx) + dfwx, ((a[1, i] + sfy) + sfwy) + dfwy
            vmax = vel_max_p if is_p[i] else vel_max_e
            nvx = min(max(vx + (fx / mass) * dt, -vmax), vmax)
            nvy = min(max(vy + (fy / mass) * dt, -vmax), vmax)
            nx, ny = x + nvx * dt, y + nvy * dt
            if periodic:
                nx = nx + 2 * hw if nx < bx_min else (nx - 2 * hw if nx > bx_max else nx)
                ny = ny + 2 * hh if ny < by_min else (ny - 2 * hh if ny > by_max else ny)
            elif billiards:
                if (nx - size_a - bx_min < 0 and nvx < 0) or (bx_max - (nx + size_a) < 0 and nvx > 0):
                    nvx = -nvx
                if (ny - size_a - by_min < 0 and nvy < 0) or (by_max - (ny + size_a) < 0 and nvy > 0):
                    nvy = -nvy
            new_p[:, i], new_dp[:, i] = (nx, ny), (nvx, nvy)
        p, dp = new_p, new_dp
    obs_dim = 4 * (2 * TOPO_NEI_MAX + (1 if self_state else 0))
    obs, rew = np.zeros((obs_dim, n)), np.zeros(n)
    nbr = np.full((n, 2 * TOPO_NEI_MAX), -1, dtype=np.int32)
    for i in range(n):
        row = 0
        if self_state:
            obs[0:4, i] = (p[0, i], p[1, i], dp[0, i], dp[1, i]); row = 4
        own = [j for j in range(n) if is_p[j] == is_p[i] and j != i]
        other = [j for j in range(n) if is_p[j] != is_p[i]]
        d_other, captures = [], 0
        for which, js in enumerate((own, other)):
            cand = []
            for j in js:
                rel = p[:, [j]] - p[:, [i]]
                if periodic:
                    rel = _wrap(rel, half)
                s = rel[0, 0] * rel[0, 0] + rel[1, 0] * rel[1, 0]
                d = np.sqrt(s)
                if which:
                    d_other.append(d); captures += int(d - 2 * size_a < 0)
                if d < d_sen:
                    cand.append((s, j))
            cand.sort()
            for q, (s, j) in enumerate(cand[:TOPO_NEI_MAX]):
                rel = p[:, [j]] - p[:, [i]]
                if periodic:
                    rel = _wrap(rel, half)
                obs[row + 4 * q: row + 4 * q + 4, i] = (rel[0, 0], rel[1, 0], dp[0, j] - dp[0, i], dp[1, j] - dp[1, i])
                nbr[i, which * TOPO_NEI_MAX + q] = j
            row += 4 * TOPO_NEI_MAX
        r = 0.0
        if other:
            dn = min(d_other)
            r = captures - 0.1 * dn if is_p[i] else -captures + 0.1 * dn
        x, y = p[0, i], p[1, i]
        if not periodic and (x - size_a - bx_min < 0 or by_max - (y + size_a) < 0 or bx_max - (x + size_a) < 0 or y - size_a - by_min < 0):
            r = r - 0.1
        rew[i] = r
    return p, dp, obs, rew, nbr
