"""Experiment: does running two half-batches on two streams (first half of one overlapping the second half of the other) beat
one full batch?  (profiling aid)  usage: python tools/exp_two_streams.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from marl_llm_b200.batched import BatchedAssemblySim, r_avoid_for
shapes = bench.load_shapes()
n_a, ngm = 30, int(shapes["n_g"].max())
r_avoid = r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"])

def make(E, seed):
    s = BatchedAssemblySim(E, n_a, ngm, r_avoid)
    s.set_shapes(shapes["grid_origin"], shapes["l_cell"])
    s.reset(seed=seed)
    a = torch.empty(8, E, 2, n_a, device="cuda")
    for r in range(8):
        s.fill_actions(a[r], seed=226, step=r)
    return s, a

def timed(fn, k=50, w=60):
    for t in range(w): fn(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(k): fn(t)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

full, af = make(65536, 1)
print("one stream, 65536 envs:", timed(lambda t: full.step(af[t % 8])), "ms/step")
del full, af
h1, a1 = make(32768, 1); h2, a2 = make(32768, 2)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both(t):
    with torch.cuda.stream(s1): h1.step(a1[t % 8])
    with torch.cuda.stream(s2): h2.step(a2[t % 8])
def both_timed(k=50, w=60):
    for t in range(w): both(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(k): both(t)
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("two streams, 2 x 32768 envs:", both_timed(), "ms/step (both halves)")
print("same two handles, one stream:", timed(lambda t: (h1.step(a1[t % 8]), h2.step(a2[t % 8]))), "ms/step")
