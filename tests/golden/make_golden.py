"""Generate tests/golden/*.npz by running the REFERENCE'S OWN code (build container only).

    python tests/golden/make_golden.py

The reference's SlateState / SeqSlateState / FeatureUtil / RecDataBase / RecEnvBase /
SlateRecEnv.obs_fn / forward run unmodified from /root/reference through the stub harness
(oracle/ref_harness.py); the TF session half is the NumPy DIEN restatement (oracle/dien_np.py).
Every fixture stores the inputs (log rows as text + SoA, catalog text, config, weight seed) and
what the reference produced at every step (features, masks, obs, reward, done, offline_*).
The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from oracle.dien_np import DienOracle  # noqa: E402
from rl4rs_b200 import synth  # noqa: E402

BASE_CFG = {"epoch": 1, "maxlen": 64, "action_size": 284, "class_num": 2, "dense_feature_num": 432,
            "category_feature_num": 21, "category_hash_size": 5000, "seq_num": 2, "emb_size": 128,
            "page_items": 9, "hidden_units": 128, "max_steps": 9, "action_emb_size": 32}

TUTORIAL_RECORD_CELL = 4   # tutorial.ipynb cell whose output holds the one real log record


def tutorial_record():
    nb = json.load(open(os.path.join(ref_harness.REFERENCE_ROOT, "tutorial.ipynb")))
    text = "".join(nb["cells"][TUTORIAL_RECORD_CELL]["outputs"][0]["text"])
    return text.split("\n")[1].strip()


def obs_to_arrays(obs, cfg):
    """Reference obs (list of dicts / ndarray) -> dict of batched arrays."""
    out = {}
    if isinstance(obs, list) and isinstance(obs[0], dict):
        for k in obs[0]:
            out[k] = np.stack([np.asarray(o[k]) for o in obs])
    else:
        out["obs"] = np.asarray(obs)
    return out


def run_reference(cfg, records, catalog_text, weights, seq, actions_fn, n_episodes=1, seed=None,
                  reset_file=False, full_stack=False, net=None):
    """Roll the reference env and record everything it exposes, per step.

    full_stack=False (how the committed fixtures were made): RecSimBase.__init__'s TensorFlow half is bypassed and the
    NumPy oracle network is plugged in as obs_layer / reward_layer (ref_harness.make_reference_env).
    full_stack=True: the reference constructs ITSELF -- SlateRecEnv.__init__ / RecSimBase.__init__, rl4rs/nets/<algo>.py's
    get_model, tf.train.Saver().restore of a Saver-format checkpoint written by rl4rs_b200.utils.tf_checkpoint -- over
    the eager layer stand-ins of oracle/tf_eager_stub.py (tests/test_reference_graph.py replays fixtures this way)."""
    tmp = tempfile.mkdtemp()
    sample_file = os.path.join(tmp, "log.csv")
    item_file = os.path.join(tmp, "item_info.csv")
    with open(sample_file, "w") as f:
        f.write("\n".join(records) + "\n")
    with open(item_file, "w") as f:
        f.write(catalog_text)
    cfg = dict(cfg, sample_file=sample_file, iteminfo_file=item_file, model_file="unused")
    if seed is not None:
        np.random.seed(seed)
    if not full_stack:
        dien = net if net is not None else DienOracle(weights, np.float32)     # any obs_layer / reward_layer pair
        env = ref_harness.make_reference_env(cfg, (dien.obs_layer, dien.reward_layer), seq=seq)
        return _roll(env, cfg, actions_fn, n_episodes, reset_file)
    from oracle import tf_eager_stub
    from rl4rs_b200.utils import tf_checkpoint
    algo = cfg.get("algo", "dien")                     # slate.py:239-242: rl4rs/nets/<algo>.py
    names = getattr(tf_checkpoint, algo + "_variable_names")(cfg)
    tensors = {names[k]: np.asarray(v, np.float32) for k, v in weights.items() if k in names}
    if algo == "dnn":                                  # nets/dnn.py also builds a sequence embedding that feeds nothing
        tensors["embedding_1/embeddings"] = np.zeros((cfg["category_hash_size"], cfg["emb_size"]), np.float32)
    cfg["model_file"] = tf_checkpoint.write_bundle(os.path.join(tmp, "simulator"), tensors)
    with tf_eager_stub.full_reference_stack(cfg) as (ref_base, ref_slate, ref_seqslate, stack):
        sim = (ref_seqslate.SeqSlateRecEnv(cfg, ref_seqslate.SeqSlateState) if seq
               else ref_slate.SlateRecEnv(cfg, ref_slate.SlateState))
        assert stack.restored_from == cfg["model_file"]
        rec = _roll(ref_base.RecEnvBase(sim), cfg, actions_fn, n_episodes, reset_file)
        assert stack.calls > 0 or cfg.get("rawstate_as_obs", False)
    return rec


def _roll(env, cfg, actions_fn, n_episodes, reset_file):
    futil = env.sim.FeatureUtil
    rec = {}

    def put(key, val):
        rec.setdefault(key, []).append(np.asarray(val))

    for ep in range(n_episodes):
        obs = env.reset(reset_file=reset_file)
        for k, v in obs_to_arrays(obs, cfg).items():
            put("reset_" + k, v)
        put("reset_user", np.asarray([int(u) for u in env.samples.user], dtype=np.int64))
        for t in range(cfg["max_steps"]):
            off_a = np.asarray(env.offline_action)
            put("offline_action", off_a)
            a = actions_fn(env, ep, t, off_a)
            put("action_in", np.asarray(a))
            obs, reward, done, info = env.step(a)
            st = env.samples.state
            raw = st["state"] if isinstance(st, dict) else st
            feat, _ = futil.feature_extraction(raw)
            put("seq", feat[0]); put("dense", feat[1]); put("cat", feat[2])
            for k, v in obs_to_arrays(obs, cfg).items():
                put("step_" + k, v)
            put("reward", np.asarray(reward, dtype=np.float64))
            put("done", np.asarray(done, dtype=np.int64))
            put("offline_reward", np.asarray(env.offline_reward, dtype=np.float64))
            put("prev_actions", env.samples.prev_actions.copy())
            put("violation", env.samples.get_violation())
            if cfg.get("simulator_info_fetch") and "click_p" in info[0]:
                put("click_p", np.stack([i["click_p"] for i in info]))
    out = {}
    for k, v in rec.items():
        out[k] = np.stack(v)
    return out


def save(name, cfg, seq, records, catalog, weight_kw, rec, extra=None):
    meta = {"config": cfg, "seq": seq, "weights": weight_kw, "records": records,
            "catalog_text": catalog.to_text()}
    meta.update(extra or {})
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **rec)
    print("wrote %s (%.1f KB): %s" % (path, os.path.getsize(path) / 1024.0,
                                      {k: v.shape for k, v in rec.items()}))


def main():
    assert ref_harness.reference_available(), "needs /root/reference"
    rs = np.random.RandomState(7)

    # --- 1. tutorial record x4, real catalog, discrete offline replay, rllib mask (Appendix C) ----
    real_cat = synth.Catalog.from_file(os.path.join(ref_harness.REFERENCE_ROOT, "dataset", "item_info.csv"))
    cfg = dict(BASE_CFG, batch_size=4, cache_size=4, is_eval=True, support_rllib_mask=True,
               category_hash_size=100000)
    wkw = {"seed": synth.WEIGHT_SEED, "stress": 1.0, "bias_noise": 0.0}
    weights = synth.make_weights(cfg, **wkw)
    recs = [tutorial_record()] * 16
    rec = run_reference(cfg, recs, real_cat.to_text(), weights, False,
                        lambda env, ep, t, off: off)
    save("tutorial_slate_rllib", cfg, False, recs, real_cat, wkw, rec,
         extra={"known": {"mask_popcounts": [38, 37, 108, 107, 106, 136, 135, 134, 1],
                          "offline_reward": 163.0}})

    cat = synth.make_catalog()
    wkw = {"seed": synth.WEIGHT_SEED, "stress": 2.0, "bias_noise": 0.1, "bounded_scores": True}

    # --- 2. Slate, synthetic rows (5% corrupted + forced), discrete replay, rllib mask, 2 episodes ----
    log = synth.make_log(96, pages=1, catalog=cat, hash_size=5000, corrupt_frac=0.25, keep_hist=True)
    recs = synth.render_records(log, cat)
    cfg = dict(BASE_CFG, batch_size=8, cache_size=8, is_eval=True, support_rllib_mask=True,
               simulator_info_fetch=True)
    weights = synth.make_weights(cfg, **wkw)
    rec = run_reference(cfg, recs, cat.to_text(), weights, False, lambda env, ep, t, off: off,
                        n_episodes=2)
    save("slate_rllib_replay", cfg, False, recs, cat, wkw, rec)

    # --- 3. Slate, mixed logged/random (often invalid) discrete actions, plain obs, train-mode sampling ----
    cfg = dict(BASE_CFG, batch_size=8, cache_size=32, is_eval=False)
    rs_a = np.random.RandomState(11)
    rec = run_reference(cfg, recs, cat.to_text(), weights, False,
                        lambda env, ep, t, off: np.where(rs_a.rand(8) < 0.85, off, rs_a.randint(0, 284, 8)),
                        n_episodes=2, seed=5)
    save("slate_plain_random", cfg, False, recs, cat, wkw, rec, extra={"np_seed": 5})

    # --- 4. Slate, d3rl mask obs, replay ----
    cfg = dict(BASE_CFG, batch_size=8, cache_size=8, is_eval=True, support_d3rl_mask=True)
    rec = run_reference(cfg, recs, cat.to_text(), weights, False, lambda env, ep, t, off: off)
    save("slate_d3rl_replay", cfg, False, recs, cat, wkw, rec)

    # --- 5. Slate, continuous actions + kNN with mask (incl. all-zero / all-one ties) ----
    cfg = dict(BASE_CFG, batch_size=8, cache_size=8, is_eval=True, support_conti_env=True,
               support_rllib_mask=True)
    rs_c = np.random.RandomState(13)

    def conti(env, ep, t, off):
        a = rs_c.uniform(-1, 1, (8, 32)).astype(np.float32)
        a[0] = 0.0                      # all-zero action: every score ties at 0 (Q18)
        a[1] = 1.0                      # tutorial.ipynb:230 all-ones action
        a[2] = off[2]                   # the logged item's own embedding (f64 -> f32)
        return a

    rec = run_reference(cfg, recs, cat.to_text(), weights, False, conti)
    save("slate_conti_knn", cfg, False, recs, cat, wkw, rec)

    # --- 6. Slate, rawstate obs + rllib mask ----
    cfg = dict(BASE_CFG, batch_size=8, cache_size=8, is_eval=True, rawstate_as_obs=True,
               support_rllib_mask=True)
    rec = run_reference(cfg, recs, cat.to_text(), weights, False, lambda env, ep, t, off: off)
    save("slate_rawstate_replay", cfg, False, recs, cat, wkw, rec)

    # --- 7. SeqSlate 36 steps, rllib mask, replay (violation zeroing on) ----
    log4 = synth.make_log(48, pages=4, catalog=cat, hash_size=5000, corrupt_frac=0.25, keep_hist=True)
    recs4 = synth.render_records(log4, cat)
    cfg = dict(BASE_CFG, batch_size=6, cache_size=6, is_eval=True, support_rllib_mask=True,
               max_steps=36)
    rec = run_reference(cfg, recs4, cat.to_text(), weights, True, lambda env, ep, t, off: off)
    save("seqslate36_rllib_replay", cfg, True, recs4, cat, wkw, rec)

    # --- 8. SeqSlate 27 steps (BASELINE config 3), plain obs (no violation zeroing), random actions ----
    cfg = dict(BASE_CFG, batch_size=6, cache_size=6, is_eval=True, max_steps=27)
    rs_b = np.random.RandomState(17)
    rec = run_reference(cfg, recs4, cat.to_text(), weights, True,
                        lambda env, ep, t, off: np.where(rs_b.rand(6) < 0.7, off, rs_b.randint(0, 284, 6)))
    save("seqslate27_plain_mixed", cfg, True, recs4, cat, wkw, rec)

    # --- 9. SeqSlate 36, d3rl mask, conti kNN ----
    cfg = dict(BASE_CFG, batch_size=6, cache_size=6, is_eval=True, support_d3rl_mask=True,
               support_conti_env=True, max_steps=36)
    rs_d = np.random.RandomState(19)
    rec = run_reference(cfg, recs4, cat.to_text(), weights, True,
                        lambda env, ep, t, off: rs_d.uniform(-1, 1, (6, 32)))
    save("seqslate36_d3rl_conti", cfg, True, recs4, cat, wkw, rec)

    # --- 10. file-cursor semantics: tiny file, wrap-around at EOF, train sampling (Q19, Q20) ----
    cfg = dict(BASE_CFG, batch_size=4, cache_size=10, is_eval=False, rawstate_as_obs=True)
    rec = run_reference(cfg, recs[:23], cat.to_text(), weights, False, lambda env, ep, t, off: off,
                        n_episodes=4, seed=3)
    save("slate_cursor_wrap", cfg, False, recs[:23], cat, wkw, rec, extra={"np_seed": 3})


if __name__ == "__main__":
    main()
