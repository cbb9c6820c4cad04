"""Import the REAL, unmodified reference environment from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by tests/golden/make_goldens.py and by the here-only tests that pin the
C restatement (oracle/assembly_oracle.c) against the reference itself.  Nothing in the product,
bench.py or the `-m gpu` tests may import this module: /root/reference does not exist on the GPU box.

What it works around:
  * envs_cplus/c_lib.py:14-21 looks for build/libAssemblyEnv.so beside itself, in a read-only tree
    -> ctypes.CDLL is patched during import to hand back oracle/_ref/libAssemblyEnv.so, which
       oracle/Makefile compiles from the reference's own AssemblyEnv.cpp with its own flags.
  * assembly.py:7-8,90 imports matplotlib (absent here) only for render() -> stub modules.
  * assembly.py:112-119 wants fig/results.pkl (missing blob) -> rebuilt from tests/golden/shapes.npz.
"""
import argparse
import ctypes
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_ROOT = os.environ.get("SWARM_REF_ROOT", "/root/reference")
REF_SO = os.path.join(HERE, "_ref", "libAssemblyEnv.so")
SHAPES_NPZ = os.path.join(REPO, "tests", "golden", "shapes.npz")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "cus_gym")) and os.path.isfile(REF_SO)


def load_shapes(path=SHAPES_NPZ):
    z = np.load(path)
    n_g = z["n_g"]
    return dict(
        l_cell=[float(v) for v in z["l_cell"]],
        grid_coords=[np.ascontiguousarray(z["grid_coords"][k, :n_g[k]]) for k in range(len(n_g))],
        shape_bound_points=[z["shape_bound_points"][k].copy() for k in range(len(n_g))],
        image_hw=z["image_hw"],
    )


def write_results_pkl(path=None):
    """The pickle layout assembly.py:116-119 reads; bitmaps replaced by 1x1 placeholders."""
    sh = load_shapes()
    blob = {
        "l_cell": sh["l_cell"],
        "grid_coords": sh["grid_coords"],
        "binary_image": [np.zeros((1, 1)) for _ in sh["l_cell"]],
        "shape_bound_points": sh["shape_bound_points"],
    }
    if path is None:
        fd, path = tempfile.mkstemp(prefix="swarm_results_", suffix=".pkl")
        os.close(fd)
    with open(path, "wb") as f:
        pickle.dump(blob, f)
    return path


_gym = None


def import_reference_gym():
    """Returns the reference's vendored `gym` package with the C++ library bound to oracle/_ref."""
    global _gym
    if _gym is not None:
        return _gym
    if not available():
        raise RuntimeError("reference checkout or oracle/_ref/libAssemblyEnv.so missing "
                           "(run `make -C oracle` in the build container)")
    mpl, plt, anim = (types.ModuleType(n) for n in
                      ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"))
    plt.figure = lambda *a, **k: None
    anim.FFMpegWriter = object
    mpl.pyplot, mpl.animation = plt, anim
    for name, mod in (("matplotlib", mpl), ("matplotlib.pyplot", plt), ("matplotlib.animation", anim)):
        sys.modules.setdefault(name, mod)

    real_cdll = ctypes.CDLL

    class _Redirect(real_cdll):
        def __init__(self, name, *a, **k):
            if isinstance(name, str) and name.endswith("libAssemblyEnv.so"):
                name = REF_SO
            super().__init__(name, *a, **k)

    sys.path.append(os.path.join(REF_ROOT, "cus_gym"))     # appended: cus_gym/tests must not shadow this repo's tests package
    ctypes.CDLL = _Redirect
    try:
        import gym  # the reference's fork (cus_gym/gym)
        from gym.envs.customized_envs import assembly  # noqa: F401  (binds _LIB now)
        import gym.wrappers  # noqa: F401
    finally:
        ctypes.CDLL = real_cdll
    assert os.path.realpath(gym.__file__).startswith(os.path.realpath(REF_ROOT)), gym.__file__
    _gym = gym
    return gym


def default_args(n_a=30, results_file=None, training_method="llm_rl", agent_strategy="input",
                 is_boundary=True, is_collected=False, is_con_self_state=True):
    """Namespace with the fields assembly.py:93-112 reads; defaults from assembly_cfg.py:152-166."""
    return argparse.Namespace(
        n_a=n_a, is_boundary=is_boundary, is_con_self_state=is_con_self_state, is_feature_norm=False,
        dynamics_mode="Cartesian", render_traj=False, traj_len=15, agent_strategy=agent_strategy,
        training_method=training_method, is_collected=is_collected,
        results_file=results_file or write_results_pkl(), video=False)


def make_env(n_a=30, **kw):
    """gym.make('AssemblySwarm-v0').unwrapped wrapped like train_assembly.py:48-50."""
    gym = import_reference_gym()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        base = gym.make("AssemblySwarm-v0").unwrapped
        env = gym.wrappers.AssemblySwarmWrapper(base, default_args(n_a=n_a, **kw))
    return env
