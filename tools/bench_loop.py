"""Device-resident rollout loop at the BASELINE config-3 shape: policy -> env.step -> buffer.push per step, CUDA-event
timed per stage and as a whole (profiling aid for DESIGN.md)."""
import json, os, sys
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from marl_llm_b200.batched import BatchedAssemblySim, r_avoid_for
from marl_llm_b200.policy import DevicePolicy
from marl_llm_b200.rollout import ReplayBufferAgent

E, n_a = int(os.environ.get("LOOP_ENVS", 65536)), 30
prec = sys.argv[1] if len(sys.argv) > 1 else "f16_tc"
shapes = bench.load_shapes()
ngm = int(shapes["n_g"].max())
sim = BatchedAssemblySim(E, n_a, ngm, r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"]))
sim.set_shapes(shapes["grid_origin"], shapes["l_cell"])
sim.reset(seed=1)
D, H, A = sim.obs_dim, 180, 2
sd = {}
torch.manual_seed(0)
for name, (o, i) in (("fc1", (H, D)), ("fc2", (H, H)), ("fc3", (H, H)), ("fc4", (A, H))):
    l = nn.Linear(i, o); sd[name + ".weight"] = l.weight; sd[name + ".bias"] = l.bias
pol = DevicePolicy(D, A, H, precision=prec, noise_scale=0.5).load_state_dict(sd)
buf = ReplayBufferAgent(8, E * n_a, slice(0, n_a), D, A)       # 8 steps of history: 8 x 1.97M rows (24 GB)
obs_prev = sim.obs.clone(); act = torch.empty(E, A, n_a, device="cuda")
idx = slice(0, n_a)
pair = [sim.obs, torch.empty_like(sim.obs)]
def one():
    prev, spare = pair
    _, lp = pol.step(prev, explore=True, out=act)
    sim.set_obs_buffer(spare)
    nxt, rew, done, _, prior = sim.step(act)
    buf.push(prev, act, rew, nxt, done, idx, prior, lp)
    pair[0], pair[1] = spare, prev
if len(sys.argv) > 2 and sys.argv[2] == "ring":
    from marl_llm_b200.episode_ring import EpisodeRing
    del buf
    ring = EpisodeRing(8, E, n_a, D, A)
    state = {"t": 0}
    def one():                                        # noqa: F811
        prev, spare = pair
        t = state["t"]
        if t == ring.T:
            ring.begin(); t = 0
        _, lp = pol.step(prev, explore=True, out=act, rows_out=ring.slot(t))
        sim.set_obs_buffer(spare)
        nxt, rew, done, _, prior = sim.step(act)
        ring.record(t, act, rew, done, prior, lp)
        state["t"] = t + 1
        pair[0], pair[1] = spare, prev
for _ in range(20): one()
torch.cuda.synchronize()
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): one()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"loop": ("policy(%s, rows_out) -> step -> small push (time-indexed ring)" if len(sys.argv) > 2 else "policy(%s) -> step -> push (obs double-buffered)") % prec, "envs": E, "n_a": n_a, "ms_per_step": ms,
                  "agent_steps_per_s": E * n_a / ms * 1e3, "mean_reward": float(sim.reward.mean())}))
