"""One small Slate episode, one SeqSlate episode (27 steps) and one episode of each other simulator (dnn, widedeep, lstm) through the env API -- the
workload for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  --log-file gpurun_out/memcheck.log  python tools/sanitize_episode.py
    compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck.log python tools/sanitize_episode.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from rl4rs_b200 import synth, gymshim  # noqa: E402
from rl4rs_b200.env.slate import SlateRecEnv, SlateState  # noqa: E402
from rl4rs_b200.env.seqslate import SeqSlateRecEnv, SeqSlateState  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160          # 2 row tiles, the second one ragged
for seq, algo in ((False, "dien"), (True, "dien"), (False, "dnn"), (False, "widedeep"), (True, "lstm")):
    cfg = {"maxlen": 64, "batch_size": B, "action_size": 284, "class_num": 2, "dense_feature_num": 432,
           "category_feature_num": 21, "category_hash_size": 5000, "seq_num": 2, "emb_size": 128, "hidden_units": 128,
           "max_steps": 27 if seq else 9, "page_items": 9, "action_emb_size": 32, "is_eval": True, "cache_size": B,
           "support_rllib_mask": True, "simulator_info_fetch": True, "output_format": "numpy", "algo": algo}
    cat = synth.make_catalog()
    log = synth.make_log(4 * B, pages=4 if seq else 1, catalog=cat, hash_size=5000)
    w = {"dnn": synth.make_dnn_weights, "widedeep": synth.make_widedeep_weights, "lstm": synth.make_lstm_weights}.get(algo, synth.make_weights)(cfg)
    c = dict(cfg, catalog=cat, log=log, weights=w)
    sim = SeqSlateRecEnv(c, state_cls=SeqSlateState) if seq else SlateRecEnv(c, state_cls=SlateState)
    env = gymshim.make("SeqSlateRecEnv-v0" if seq else "SlateRecEnv-v0", recsim=sim)
    obs = env.reset()
    tot = 0.0
    for t in range(cfg["max_steps"]):
        obs, reward, done, info = env.step(env.offline_action)
        tot += float(np.sum(reward))
    print("%s %s: B=%d episode done, sum reward %.3f, launches %d" % ("SeqSlate" if seq else "Slate", algo, B, tot,
                                                                          sim.engine.launch_count()))
    sim.engine.close()
