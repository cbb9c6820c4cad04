"""The reference's CPU step path, runnable where /root/reference is NOT mounted (the GPU box).

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py --impl reference and its cpu_baseline leg, tests).

The heavy lifting is the reference's OWN C++ (oracle/_ref/libAssemblyEnv.so, compiled by oracle/Makefile from
/root/reference/.../AssemblyEnv.cpp with the reference's flags — a build artefact that travels with the repo
snapshot).  What cannot travel is the reference's Python file, so the NumPy glue that assembly.py wraps around the
five C calls is restated here, operation for operation, with the same temporaries (np.tile pair matrices, 2x4 wall
matrices, ...) so that the timing is representative of what the reference executes per step:
    reset  : assembly.py:156-223      step : assembly.py:487-666      pair distances : assembly.py:442-457
tests/test_oracle_vs_reference.py::test_ref_glue_* pins it bit-exactly against the real class in the build container.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libAssemblyEnv.so")


def available():
    return os.path.isfile(REF_SO)


def _d(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _b(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_bool))


class RefEnv:
    """One reference env (agent_strategy='input', Cartesian, is_boundary=True, training_method='llm_rl')."""

    def __init__(self, n_a, shapes, lib=None):
        self.lib = lib or ctypes.CDLL(REF_SO)
        self.n_a, self.dim = n_a, 2
        self.l_cells = list(shapes["l_cell"])
        self.grid_origins = shapes["grid_origin"]                       # list of [2, n_g]
        self.n_gs = [g.shape[1] for g in self.grid_origins]
        self.r_avoid = round(np.sqrt(4 * np.min(self.n_gs) / (n_a * np.pi)) * np.min(self.l_cells), 2)   # :124
        self.topo_nei_max, self.num_obs_grid_max, self.num_occupied_grid_max = 6, 80, 200
        self.obs_dim_agent = 2 * 2 * (6 + 1 + 1) + 2 * 80                # :801
        self.size_a, self.k_ball, self.k_wall, self.c_wall = 0.035, 30, 100, 5
        self.dt, self.Vel_max, self.half = 0.1, 0.8, 2.4
        self.size = np.array([self.size_a for _ in range(n_a)])           # :782-788
        sizes = np.tile(self.size.reshape(n_a, 1), (1, n_a))
        sizes = sizes + sizes.T
        sizes[np.arange(n_a), np.arange(n_a)] = 0
        self.sizes = sizes
        self.m = np.array([1 for _ in range(n_a)])                        # :790-793
        self.is_collide_b2w = np.zeros((4, n_a), dtype=bool)
        self.d_b2w = np.ones((4, n_a))
        self.heading = np.zeros((2, n_a))

    def reset(self):
        """assembly.py:156-223, same NumPy-global RNG call order."""
        k = np.random.randint(0, len(self.l_cells))
        self.l_cell = self.l_cells[k]
        origin = self.grid_origins[k]
        ang = np.pi * np.random.uniform(-1, 1)
        R = np.array([[np.cos(ang), np.sin(ang)], [-np.sin(ang), np.cos(ang)]])
        origin = np.dot(R, origin)
        self.n_g = origin.shape[1]
        np.random.uniform(-1.2, 1.2, (2, 1))                              # drawn and discarded (:182)
        off = np.array([[np.random.uniform(-self.half + 1, self.half - 1),
                         np.random.uniform(-self.half + 1, self.half - 1)]]).T
        self.grid_center = origin.copy() + off
        self.boundary_pos = np.array([-self.half, self.half, self.half, -self.half], dtype=np.float64)
        self.d_sen = 0.4
        if np.random.uniform(-1, 1) > 0:
            self.p = np.concatenate((np.random.uniform(-self.half, self.half, (1, self.n_a)),
                                     np.random.uniform(-self.half, self.half, (1, self.n_a))), axis=0)
        else:
            self.p = np.random.uniform(-1, 1, (2, self.n_a)) + np.array(
                [[np.random.uniform(-self.half + 1, self.half - 1), np.random.uniform(-self.half + 1, self.half - 1)]]).T
        self.dp = np.random.uniform(-0.5, 0.5, (2, self.n_a))
        return self._get_obs()

    def _get_obs(self):
        """assembly.py:225-255"""
        n = self.n_a
        self.obs = np.zeros((self.obs_dim_agent, n))
        self.neighbor_index = -1 * np.ones((n, self.topo_nei_max), dtype=np.int32)
        self.in_flags = np.zeros(n, dtype=np.int32)
        self.sensed_index = -1 * np.ones((n, self.num_obs_grid_max), dtype=np.int32)
        self.occupied_index = -1 * np.ones((n, self.num_occupied_grid_max), dtype=np.int32)
        cond = np.array([False, True, True, False])
        self.lib._get_observation(
            _d(self.p), _d(self.dp), _d(self.heading), _d(self.obs), _d(self.boundary_pos), _d(self.grid_center),
            _i(self.neighbor_index), _i(self.in_flags), _i(self.sensed_index), _i(self.occupied_index),
            ctypes.c_double(self.d_sen), ctypes.c_double(self.r_avoid), ctypes.c_double(self.l_cell),
            ctypes.c_double(self.Vel_max), ctypes.c_int(self.topo_nei_max), ctypes.c_int(self.num_obs_grid_max),
            ctypes.c_int(self.num_occupied_grid_max), ctypes.c_int(n), ctypes.c_int(self.n_g),
            ctypes.c_int(self.obs_dim_agent), ctypes.c_int(2), _b(cond))
        return self.obs

    def step(self, a):
        """assembly.py:487-666"""
        n = self.n_a
        # pair distances, :442-457
        all_pos = np.tile(self.p, (n, 1))
        my_pos = np.tile(self.p.T.reshape(2 * n, 1), (1, n))
        rel = all_pos - my_pos
        d_center = np.sqrt(rel[::2, :] ** 2 + rel[1::2, :] ** 2)
        d_edge = d_center - self.sizes
        collide = (d_edge < 0)
        d_edge = np.abs(d_edge)
        sf_b2b = np.zeros((2, n))
        self.lib._sf_b2b_all(_d(self.p), _d(sf_b2b), _d(d_edge), _b(collide), _d(self.boundary_pos), _d(d_center),
                             ctypes.c_int(n), ctypes.c_int(2), ctypes.c_double(self.k_ball), ctypes.c_bool(False))
        # walls, :515-518
        self.lib._get_dist_b2w(_d(self.p), _d(self.size), _d(self.d_b2w), _b(self.is_collide_b2w), ctypes.c_int(2),
                               ctypes.c_int(n), _d(self.boundary_pos))
        sf_b2w = np.array([[1, 0, -1, 0], [0, -1, 0, 1]]).dot(self.is_collide_b2w * self.d_b2w) * self.k_wall
        df_b2w = np.array([[-1, 0, -1, 0], [0, -1, 0, -1]]).dot(
            self.is_collide_b2w * np.concatenate((self.dp, self.dp), axis=0)) * self.c_wall
        # prior, :605-624
        a_prior = np.zeros((2, n))
        self.lib.calculateActionPrior(_d(self.p), _d(self.dp), _d(a_prior), _d(self.grid_center), _i(self.neighbor_index),
                                      ctypes.c_double(self.d_sen), ctypes.c_double(self.r_avoid), ctypes.c_double(self.l_cell),
                                      ctypes.c_int(self.topo_nei_max), ctypes.c_int(n), ctypes.c_int(self.n_g), ctypes.c_int(2))
        # integrator, :631-650
        F = 1 * a + sf_b2b + sf_b2w + df_b2w
        ddp = F / self.m
        self.dp += ddp * self.dt
        self.dp = np.clip(self.dp, -self.Vel_max, self.Vel_max)
        self.p += self.dp * self.dt
        obs = self._get_obs()
        # reward, :351-380
        rew = np.zeros((1, n))
        coef = np.array([0.05], dtype=np.float64)
        cond = np.array([False, True, True, True, True], dtype=bool)
        self.lib._get_reward(_d(self.p), _d(self.dp), _d(self.heading), _d(a.astype(np.float64)), _d(rew), _d(self.boundary_pos),
                             _d(self.grid_center), _i(self.neighbor_index), _i(self.in_flags), _i(self.sensed_index),
                             _i(self.occupied_index), ctypes.c_double(self.d_sen), ctypes.c_double(self.r_avoid),
                             ctypes.c_double(self.l_cell), ctypes.c_int(self.topo_nei_max), ctypes.c_int(self.num_obs_grid_max),
                             ctypes.c_int(self.num_occupied_grid_max), ctypes.c_int(n), ctypes.c_int(self.n_g), ctypes.c_int(2),
                             _b(cond), _b(collide), _b(self.is_collide_b2w), _d(coef))
        done = np.zeros((1, n)).astype(bool)                              # :480-482
        info = np.array([None, None, None]).reshape(3, 1)                 # :484-485
        return obs, rew, done, info, a_prior


def _worker(args):
    """One process = one env looping `steps` steps `episodes` times; returns (agent_steps, seconds in env.step)."""
    import time
    n_a, shapes, seed, episodes, steps = args
    env = RefEnv(n_a, shapes)
    np.random.seed(seed)
    rng = np.random.RandomState(seed + 1)
    acts = rng.uniform(-1, 1, (steps, 2, n_a)).astype(np.float32)         # pre-generated, outside the timed region
    spent = 0.0
    for _ in range(episodes):
        env.reset()
        t0 = time.perf_counter()
        for t in range(steps):
            env.step(acts[t])
        spent += time.perf_counter() - t0
    return episodes * steps * n_a, spent


def timed_rollouts(n_a, shapes, procs, episodes, steps, seed=226):
    """`procs` independent processes (the reference is single-threaded and has no vector env),
    each stepping its own env.  Returns aggregate agent-steps/s = sum(agent_steps) / max(process time)."""
    import multiprocessing as mp
    import time
    ctx = mp.get_context("fork")
    jobs = [(n_a, shapes, seed + 17 * k, episodes, steps) for k in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_worker(jobs[0])]
    else:
        with ctx.Pool(procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return dict(agent_steps=total, seconds=slowest, wall=wall, value=total / slowest)
