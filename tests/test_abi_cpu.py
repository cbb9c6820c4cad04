"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/swarm_b200.h declares, the ctypes
mirrors match the C structs, host-only logic is right, and it FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from marl_llm_b200 import _lib
from marl_llm_b200.build import build_library

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.load()


def declared_symbols():
    hdr = open(os.path.join(REPO, "include", "swarm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr)
    return sorted({n for n in names if n.startswith(("swarm_", "_get_", "_sf_")) or n == "calculateActionPrior"})


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert set(_lib.LEGACY_SYMBOLS) <= set(syms) and len(syms) >= 20
    for s in syms:
        assert getattr(lib, s) is not None, s
    assert set(syms) == set(_lib.LEGACY_SYMBOLS + _lib.BATCHED_SYMBOLS + _lib.ROLLOUT_SYMBOLS + _lib.POLICY_SYMBOLS)


def test_legacy_alias_name_exists():
    # the reference loader asks for lib{env_name}Env.so (c_lib.py:14-21)
    alias = os.path.join(os.path.dirname(_lib.LIB_PATH), "libAssemblyEnv.so")
    assert os.path.exists(alias)
    h = C.CDLL(alias)
    for s in _lib.LEGACY_SYMBOLS:
        getattr(h, s)


def test_struct_sizes_and_abi_version(lib):
    assert lib.swarm_abi_version() == 2
    cfg = _lib.SwarmConfig()
    cfg.struct_size = C.sizeof(_lib.SwarmConfig)
    cfg.is_con_self_state, cfg.num_obs_grid_max = 1, 80
    assert lib.swarm_obs_dim(C.byref(cfg)) == 192          # assembly.py:801
    cfg.is_con_self_state = 0
    assert lib.swarm_obs_dim(C.byref(cfg)) == 188
    assert lib.swarm_grid_pad(536) == 544 and lib.swarm_grid_pad(32) == 32 and lib.swarm_grid_pad(1) == 32


@pytest.mark.parametrize("d", [0.4, 0.26, 0.13, 0.53, 0.07, 0.05, math.sqrt(2) * 0.06111 / 2, 1e-3, 3.0])
def test_sqrt_thresholds_are_exact(lib, d):
    """sqrt(s) < d  <=>  s < T  and  sqrt(s) <= d  <=>  s <= U, checked on the doubles around the boundary."""
    T = lib.swarm_sqrt_threshold(d, 0)
    U = lib.swarm_sqrt_threshold(d, 1)
    s = T
    for _ in range(6):
        s = np.nextafter(s, -np.inf)
    for _ in range(13):
        assert (math.sqrt(s) < d) == (s < T)
        s = np.nextafter(s, np.inf)
    s = U
    for _ in range(6):
        s = np.nextafter(s, -np.inf)
    for _ in range(13):
        assert (math.sqrt(s) <= d) == (s <= U)
        s = np.nextafter(s, np.inf)


def test_create_validates_and_never_falls_back(lib):
    import torch
    cfg = _lib.SwarmConfig()
    buf = _lib.SwarmBuffers()
    h = C.c_void_p()
    assert lib.swarm_create(C.byref(cfg), C.byref(buf), C.byref(h)) == _lib.SWARM_ERR_INVALID   # struct_size unset
    cfg.struct_size, buf.struct_size = C.sizeof(_lib.SwarmConfig), C.sizeof(_lib.SwarmBuffers)
    cfg.num_envs, cfg.n_a, cfg.n_g_max, cfg.topo_nei_max = 4, 30, 536, 5
    cfg.num_obs_grid_max, cfg.num_occupied_grid_max = 80, 200
    assert lib.swarm_create(C.byref(cfg), C.byref(buf), C.byref(h)) == _lib.SWARM_ERR_UNSUPPORTED  # topo != 6
    cfg.topo_nei_max = 6
    cfg.n_a = 5000
    assert lib.swarm_create(C.byref(cfg), C.byref(buf), C.byref(h)) == _lib.SWARM_ERR_UNSUPPORTED
    cfg.n_a = 30
    assert lib.swarm_create(C.byref(cfg), C.byref(buf), C.byref(h)) == _lib.SWARM_ERR_INVALID   # NULL buffers
    assert b"NULL" in lib.swarm_last_error()
    if not torch.cuda.is_available():
        dummy = (C.c_double * 8)()
        for name, _ in _lib.SwarmBuffers._fields_[2:]:
            if name == "a_prior":
                buf.a_prior[0] = buf.a_prior[1] = C.addressof(dummy)
            else:
                setattr(buf, name, C.addressof(dummy))
        rc = lib.swarm_create(C.byref(cfg), C.byref(buf), C.byref(h))
        assert rc in (_lib.SWARM_ERR_NO_DEVICE, _lib.SWARM_ERR_CUDA), rc
        assert h.value is None
        from marl_llm_b200.batched import BatchedAssemblySim
        with pytest.raises(_lib.SwarmError):
            BatchedAssemblySim(4, 30, 536, 0.26)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "marl_llm_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_rollout_and_policy_entry_points_fail_loudly_without_gpu(lib):
    """Groups 3 and 4 of the ABI: argument validation happens before any device access; on a box without a GPU the
    policy constructor reports SWARM_ERR_NO_DEVICE (there is no CPU implementation behind it)."""
    import torch
    b = _lib.SwarmRolloutBuffers()
    b.struct_size = 0
    rc = lib.swarm_rollout_push(C.byref(b), 0, 1, 1, 0, 1, None, None, None, None, None, _lib.SWARM_F32, None, _lib.SWARM_F32, None, None)
    assert rc == _lib.SWARM_ERR_INVALID and b"struct_size" in lib.swarm_last_error()
    b.struct_size = C.sizeof(_lib.SwarmRolloutBuffers)
    b.obs_dim, b.act_dim, b.capacity = 192, 2, 100
    rc = lib.swarm_rollout_push(C.byref(b), 0, 1, 1, 0, 1, None, None, None, None, None, _lib.SWARM_F32, None, _lib.SWARM_F32, None, None)
    assert rc == _lib.SWARM_ERR_INVALID and b"NULL" in lib.swarm_last_error()
    rc = lib.swarm_rollout_gather(C.byref(b), None, 4, None, None, None, None, None, None, None, None)
    assert rc == _lib.SWARM_ERR_INVALID
    h = C.c_void_p()
    assert lib.swarm_policy_create(0, 500, 180, 2, C.byref(h)) == _lib.SWARM_ERR_UNSUPPORTED        # obs_dim > 192
    assert lib.swarm_policy_create(0, 192, 180, 9, C.byref(h)) == _lib.SWARM_ERR_UNSUPPORTED        # act_dim > 8
    if not torch.cuda.is_available():
        assert lib.swarm_policy_create(0, 192, 180, 2, C.byref(h)) == _lib.SWARM_ERR_NO_DEVICE
        assert b"no CPU fallback" in lib.swarm_last_error()
        from marl_llm_b200.policy import DevicePolicy
        from marl_llm_b200.rollout import ReplayBufferAgent
        with pytest.raises(_lib.SwarmError):
            DevicePolicy(192, 2)
        with pytest.raises(_lib.SwarmError):
            ReplayBufferAgent(10, 30, slice(0, 30), 192, 2)
    assert lib.swarm_policy_step(None, None, 1, 1, None, None, 0, 0.1, 0, 0, None) == _lib.SWARM_ERR_INVALID
    assert lib.swarm_policy_set_precision(None, 0) == _lib.SWARM_ERR_INVALID


def test_rollout_struct_mirror_matches_header():
    hdr = open(os.path.join(REPO, "include", "swarm_b200.h")).read()
    body = re.search(r"typedef struct swarm_rollout_buffers \{(.*?)\} swarm_rollout_buffers;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for decl in body.split(";") for n in re.findall(r"\*?\s*([a-z_0-9]+)\s*(?:,|$)", decl.split(None, 1)[1] if decl.strip() else "")]
    assert names == [f[0] for f in _lib.SwarmRolloutBuffers._fields_], names
