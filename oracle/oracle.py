"""ctypes front-end for oracle/liboracle.so (the plain-C restatement in assembly_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of assembly_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, nowhere else.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")


class OrcParams(C.Structure):
    """Mirror of `orc_params` in assembly_oracle.c (field comments there cite the reference)."""
    _fields_ = [
        ("n_a", C.c_int32), ("n_g", C.c_int32), ("topo_nei_max", C.c_int32),
        ("num_obs_grid_max", C.c_int32), ("num_occupied_grid_max", C.c_int32), ("obs_dim", C.c_int32),
        ("is_con_self_state", C.c_int32), ("is_periodic", C.c_int32), ("want_prior", C.c_int32),
        ("pad_", C.c_int32),
        ("d_sen", C.c_double), ("r_avoid", C.c_double), ("l_cell", C.c_double), ("size_a", C.c_double),
        ("k_ball", C.c_double), ("k_wall", C.c_double), ("c_wall", C.c_double),
        ("dt", C.c_double), ("vel_max", C.c_double), ("mass", C.c_double),
        ("boundary_pos", C.c_double * 4),
    ]


def build(force=False):
    if force or not os.path.isfile(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "assembly_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        assert _lib.orc_params_size() == C.sizeof(OrcParams)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def r_avoid_for(n_a, n_gs, l_cells):
    """assembly.py:124"""
    return round(float(np.sqrt(4 * np.min(n_gs) / (n_a * np.pi)) * np.min(l_cells)), 2)


def make_params(n_a, n_g, l_cell, r_avoid, d_sen=0.4, is_con_self_state=True, is_periodic=False,
                want_prior=True, half_w=2.4, half_h=2.4):
    """Constants from assembly.py:27-81,128-130,193-199."""
    P = OrcParams()
    P.n_a, P.n_g = n_a, n_g
    P.topo_nei_max, P.num_obs_grid_max, P.num_occupied_grid_max = 6, 80, 200
    self_flag = 1 if is_con_self_state else 0
    P.obs_dim = 2 * 2 * (6 + 1 + self_flag) + 2 * 80                       # assembly.py:801
    P.is_con_self_state, P.is_periodic, P.want_prior = int(is_con_self_state), int(is_periodic), int(want_prior)
    P.d_sen, P.r_avoid, P.l_cell, P.size_a = d_sen, r_avoid, l_cell, 0.035
    P.k_ball, P.k_wall, P.c_wall = 30.0, 100.0, 5.0
    P.dt, P.vel_max, P.mass = 0.1, 0.8, 1.0
    P.boundary_pos[:] = [-half_w, half_h, half_w, -half_h]
    return P


class OracleBatch:
    """E independent assembly envs stepped by the C restatement.  Same per-env layouts as the
    reference; arrays are [E, ...] stacks of them.  Grid blocks are [E, 2, ng_max] (row-major,
    each env's [2, n_g] matrix stored contiguously at the start of its block)."""

    def __init__(self, params_list, nthreads=1, ng_max=None):
        self.E = len(params_list)
        self.params = (OrcParams * self.E)(*params_list)
        P0 = params_list[0]
        self.n_a, self.obs_dim = P0.n_a, P0.obs_dim
        self.ng_max = max(max(p.n_g for p in params_list), ng_max or 0)
        E, n = self.E, self.n_a
        self.p = np.zeros((E, 2, n))
        self.dp = np.zeros((E, 2, n))
        self.grid = np.zeros((E, 2 * self.ng_max))
        self.obs = np.zeros((E, self.obs_dim, n))
        self.reward = np.zeros((E, 1, n))
        self.a_prior = np.zeros((E, 2, n))
        self.neighbor_index = -np.ones((E, n, 6), dtype=np.int32)
        self.in_flags = np.zeros((E, n), dtype=np.int32)
        self.sensed_index = -np.ones((E, n, 80), dtype=np.int32)
        self.occupied_index = -np.ones((E, n, 200), dtype=np.int32)
        self.nthreads = nthreads

    def set_grid(self, e, grid_2xng):
        g = np.ascontiguousarray(grid_2xng, dtype=np.float64)
        assert g.shape == (2, self.params[e].n_g)
        self.grid[e, :g.size] = g.reshape(-1)

    def grid_of(self, e):
        ng = self.params[e].n_g
        return self.grid[e, :2 * ng].reshape(2, ng)

    def observe(self, with_reward=False):
        lib().orc_observe_batch(
            C.c_int(self.E), self.params, _dp(self.p), _dp(self.dp), _dp(self.grid),
            C.c_long(2 * self.ng_max), _dp(self.obs), _dp(self.reward) if with_reward else None,
            _ip(self.neighbor_index), _ip(self.in_flags), _ip(self.sensed_index), _ip(self.occupied_index),
            C.c_int(self.nthreads))
        return self.obs

    def step(self, act):
        act = np.ascontiguousarray(act, dtype=np.float32)
        assert act.shape == (self.E, 2, self.n_a)
        lib().orc_step_batch(
            C.c_int(self.E), self.params, _dp(self.p), _dp(self.dp), _fp(act), _dp(self.grid),
            C.c_long(2 * self.ng_max), _dp(self.obs), _dp(self.reward), _dp(self.a_prior),
            _ip(self.neighbor_index), _ip(self.in_flags), _ip(self.sensed_index), _ip(self.occupied_index),
            C.c_int(self.nthreads))
        return self.obs, self.reward, self.a_prior


def strategy_actions(ob, kind):
    """Actions of the reference's host strategies for the CURRENT state of an OracleBatch: kind 'rule' (assembly.py:530-601)
    or 'llm' (assembly.py:524-529, 876-941; uses the neighbour list of the last observation).  [E, 2, n_a] float64."""
    out = np.zeros((ob.E, 2, ob.n_a))
    for e in range(ob.E):
        P = ob.params[e]
        g = np.ascontiguousarray(ob.grid_of(e))
        if kind == "rule":
            lib().orc_rule_actions(C.byref(P), _dp(ob.p[e]), _dp(ob.dp[e]), _dp(g), _dp(out[e]))
        elif kind == "llm":
            lib().orc_llm_actions(C.byref(P), _dp(ob.p[e]), _dp(ob.dp[e]), _dp(g), _ip(ob.neighbor_index[e]), _dp(out[e]))
        else:
            raise ValueError(kind)
    return out


def fill_actions(E, n_a, seed, step, env0=0):
    """U(-1,1) float32 [E,2,n_a] from the counter-based generator shared with the CUDA side."""
    act = np.empty((E, 2, n_a), dtype=np.float32)
    lib().orc_fill_actions(C.c_int(E), C.c_int(n_a), C.c_uint64(seed), C.c_uint64(step), C.c_uint64(env0), _fp(act))
    return act


def reset_uniform(seed, episode, env, k):
    """The k-th U[0,1) draw of env `env` in swarm_reset's counter-based generator (marl_llm_b200/csrc/swarm_kernels.cuh: mix64 /
    u01), restated with NumPy uint64 arithmetic; env and k broadcast.  Lets a test rebuild the reset state from the draws."""
    with np.errstate(over="ignore"):
        u = np.uint64
        z = (u(seed) * u(0x9E3779B97F4A7C15) + u(episode) * u(0xBF58476D1CE4E5B9)
             + np.asarray(env, dtype=np.uint64) * u(0x94D049BB133111EB) + np.asarray(k, dtype=np.uint64) * u(0xD6E8FEB86659FD93))
        z = z ^ (z >> u(30)); z = z * u(0xBF58476D1CE4E5B9)
        z = z ^ (z >> u(27)); z = z * u(0x94D049BB133111EB)
        z = z ^ (z >> u(31))
    return (z >> u(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
