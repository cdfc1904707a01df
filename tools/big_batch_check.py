"""BASELINE config 3 sanity: SeqSlateRecEnv-v0, batch 16384, A2C, 27 steps -- sizes, index widths, page caches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from rl4rs_b200 import synth, gymshim
from rl4rs_b200.env.seqslate import SeqSlateRecEnv, SeqSlateState
from rl4rs_b200.trainer import get_rl_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = {"maxlen": 64, "batch_size": B, "action_size": 284, "class_num": 2, "dense_feature_num": 432,
       "category_feature_num": 21, "category_hash_size": 100000, "seq_num": 2, "emb_size": 128, "hidden_units": 128,
       "max_steps": 27, "page_items": 9, "action_emb_size": 32, "is_eval": False, "cache_size": 2048,
       "support_rllib_mask": True, "output_format": "torch"}
cat = synth.make_catalog(); log = synth.make_log(4 * B, pages=4, catalog=cat); w = synth.make_weights(cfg)
sim = SeqSlateRecEnv(dict(cfg, catalog=cat, log=log, weights=w), state_cls=SeqSlateState)
env = gymshim.make("SeqSlateRecEnv-v0", recsim=sim)
tr = get_rl_model("A2C", {}, env=env)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    r = tr.train()
    torch.cuda.synchronize(); dt = time.time() - t0
    print("iter %d: %.1f ms, %.0f transitions/s, reward %.3f, loss %.3g, mem %.1f GB" % (
        i, dt * 1e3, B * 27 / dt, r["episode_reward_mean"], r["total_loss"], torch.cuda.max_memory_allocated() / 2**30))
assert np.isfinite(r["total_loss"]) and r["episode_reward_mean"] > 0
free, total = torch.cuda.mem_get_info(); print("device memory in use %.1f GB" % ((total - free) / 2**30))
