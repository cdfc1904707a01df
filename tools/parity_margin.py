"""Print the worst (error / 1e-4 bound) of obs / reward / click_p over all golden fixtures (GPU)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import Golden, golden_names
from test_gpu_parity import make_env, obs_arrays


def ratio(got, ref, rtol=1e-4):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    scale = np.sqrt((ref ** 2).mean(axis=-1, keepdims=True)) if ref.ndim >= 2 else np.abs(ref)
    bound = rtol * np.maximum(np.abs(ref), scale) + 1e-12
    return float((np.abs(got - ref) / bound).max())


for name in golden_names():
    g = Golden(name)
    if g.config.get("rawstate_as_obs"):
        continue
    if "np_seed" in g.meta:
        np.random.seed(g.meta["np_seed"])
    env = make_env(g.config, g.seq, g.catalog, g.log, g.weights, output_format="numpy")
    w = {"obs": 0.0, "reward": 0.0, "click_p": 0.0}
    k = 0
    for ep in range(g.n_episodes):
        o = obs_arrays(env.reset())
        w["obs"] = max(w["obs"], ratio(o["obs"][:, :256], g.arr["reset_obs"][ep][:, :256]))
        for t in range(g.config["max_steps"]):
            o, r, d, info = env.step(g.arr["action_in"][k])
            o = obs_arrays(o)
            w["obs"] = max(w["obs"], ratio(o["obs"][:, :256], g.arr["step_obs"][k][:, :256]))
            w["reward"] = max(w["reward"], ratio(r, g.arr["reward"][k]))
            if "click_p" in g.arr and t == g.config["max_steps"] - 1:
                w["click_p"] = max(w["click_p"], ratio(info["click_p"], g.arr["click_p"][ep]))
            k += 1
    print("%-28s worst err/bound: obs %.3f  reward %.3f  click_p %.3f" % (name, w["obs"], w["reward"], w["click_p"]))
