"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's `dnn` user-response simulator
(config['algo'] = 'dnn', rl4rs/env/slate.py:239-242 -> rl4rs/nets/dnn.py:8-45).

PARITY UNPINNED for the floating-point network, for the same reason as oracle/dien_np.py: the reference holds no
golden vector for it and TensorFlow 1.15 (Embedding / GlobalAveragePooling1D / Dense / ELU, call sites
rl4rs/nets/utils.py:7-14,48-54 and rl4rs/nets/dnn.py:34-36) is absent.  The graph is plain Keras, restated here:

  category_feature = GlobalAveragePooling1D(Embedding(cat))            utils.py:7-14     (B, 128)
  dense_feature    = ELU(ELU(dense W1 + b1) W2 + b2)                   utils.py:48-54    (B, 128)
  sequence_feature = sequence_input_concat(seq)                        dnn.py:33 -- computed, NEVER used (dnn.py:34
                     concatenates category and dense only), so it is not evaluated here
  all   = Dense(256, ELU)(concat[category_feature, dense_feature])     dnn.py:34         (the unnamed 'dense_2')
  obs   = Dense(256, ELU, name='simulator_obs')(all)                   dnn.py:35
  probs = Dense(2, softmax, name='simulator_reward')(obs)              dnn.py:36
"""
import numpy as np


def weight_shapes(cfg=None):
    cfg = cfg or {}
    H, E, U = cfg.get("category_hash_size", 100000), cfg.get("emb_size", 128), cfg.get("hidden_units", 128)
    D = cfg.get("dense_feature_num", 432)
    return {"emb_cat": (H, E), "dense_w1": (D, U), "dense_b1": (U,), "dense_w2": (U, U), "dense_b2": (U,),
            "fc_w": (E + U, 256), "fc_b": (256,), "obs_w": (256, 256), "obs_b": (256,), "rew_w": (256, 2), "rew_b": (2,)}


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


class DnnOracle:
    """forward(seq (ignored), dense f[B,432], cat int[B,21]) -> (obs [B,256], probs [B,2])."""

    def __init__(self, weights, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.w = {k: np.asarray(v, dtype=self.dt) for k, v in weights.items()}      # no copy when the dtype already matches

    def forward(self, seq, dense, cat):
        w = self.w
        cat = np.asarray(cat).astype(np.int64)
        c = w["emb_cat"][cat].mean(axis=1).astype(self.dt)                       # utils.py:11-13
        x = _elu(np.asarray(dense).astype(self.dt) @ w["dense_w1"] + w["dense_b1"])
        x = _elu(x @ w["dense_w2"] + w["dense_b2"]).astype(self.dt)              # utils.py:50-53 (dropout inert)
        a = _elu(np.concatenate([c, x], axis=-1) @ w["fc_w"] + w["fc_b"]).astype(self.dt)     # dnn.py:34
        obs = _elu(a @ w["obs_w"] + w["obs_b"]).astype(self.dt)                  # dnn.py:35
        z = obs @ w["rew_w"] + w["rew_b"]                                        # dnn.py:36
        z = z - z.max(axis=-1, keepdims=True)
        p = np.exp(z)
        return obs, (p / p.sum(axis=-1, keepdims=True)).astype(self.dt)

    def obs_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[0].astype(np.float32)

    def reward_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[1].astype(np.float32)
