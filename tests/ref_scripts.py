"""Harness that runs the reference's own, byte-identical scripts (staged under baseline/_ref by tools/stage_reference.py)
in a subprocess — test infrastructure for the "runs unchanged" acceptance tests (north star).

Two ways of putting the B200 simulator under them:
  'legacy'  the reference's Gym fork and its real assembly.py (baseline/_ref/cus_gym) with libswarm_b200.so installed as
            envs_cplus/build/libAssemblyEnv.so — exactly INTEGRATION.md §A (c_lib.py:11-22 loads it, five ctypes symbols);
  'dropin'  marl_llm_b200/compat on PYTHONPATH instead of cus_gym: `gym.make('AssemblySwarm-v0')` returns the drop-in class
            on top of the batched C ABI.
What is arranged around the scripts (never inside them): matplotlib / tensorboardX shims (absent from this image), a working
directory that contains 'Your/Image/Folder/Path' (the placeholder assembly_cfg.py:140 ships with) holding the seven PNGs."""
import hashlib
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
SHIMS = os.path.join(REPO, "tests", "shims")
BUILD_DIR = os.path.join(REF, "cus_gym", "gym", "envs", "customized_envs", "envs_cplus", "build")


def staged():
    return os.path.isfile(os.path.join(REF, "MANIFEST.json"))


def verify_manifest(rels):
    """The staged files really are the reference's bytes (digest recorded when they were copied from /root/reference)."""
    man = json.load(open(os.path.join(REF, "MANIFEST.json")))
    for rel in rels:
        assert hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest() == man[rel], f"{rel} was modified"


def install_library(lib_path):
    """drop the shared library where c_lib.py:14-21 looks for it."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    dst = os.path.join(BUILD_DIR, "libAssemblyEnv.so")
    if os.path.lexists(dst):
        os.remove(dst)
    shutil.copyfile(lib_path, dst)
    return dst


def make_workdir(tmp):
    img = os.path.join(tmp, "Your", "Image", "Folder", "Path")
    os.makedirs(img, exist_ok=True)
    for f in os.listdir(os.path.join(REF, "fig")):
        if f.endswith(".png"):
            shutil.copyfile(os.path.join(REF, "fig", f), os.path.join(img, f))
    return tmp


def pythonpath(mode):
    gym_root = os.path.join(REF, "cus_gym") if mode == "legacy" else os.path.join(REPO, "marl_llm_b200", "compat")
    return os.pathsep.join([SHIMS, gym_root, os.path.join(REF, "marl_llm"), REPO])


def run_script(rel_script, mode, cwd, argv=(), timeout=900, extra_env=None):
    env = dict(os.environ)
    env["PYTHONPATH"] = pythonpath(mode)
    env["MPLBACKEND"] = "Agg"
    env.update(extra_env or {})
    cmd = [sys.executable, os.path.join(REF, rel_script)] + list(argv)
    r = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    return r
