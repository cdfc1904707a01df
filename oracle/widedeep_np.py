"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's `widedeep` user-response simulator
(config['algo'] = 'widedeep', rl4rs/env/slate.py:239-242 -> rl4rs/nets/widedeep.py:8-45).

PARITY UNPINNED for the floating-point network (no reference-held vector; TensorFlow 1.15 absent) -- plain Keras layers:

  category_feature = Flatten(Embedding_0(cat))                               utils.py:38-45    (B, 21*128)
  dense_feature    = ELU(ELU(dense W1 + b1) W2 + b2)                         utils.py:48-54    (B, 128)
  sequence_feature = concat_i GlobalAveragePooling1D(Embedding_1(seq[:, i])) utils.py:56-77    (B, 2*128)
                     ONE embedding layer shared by both sequences; no mask: the pad id 0 is embedded and averaged
  seq_dnn          = Dense(256, ELU)(sequence_feature)                       widedeep.py:34
  obs = 'simulator_obs' = Concatenate([seq_dnn, dense_feature, category_feature])   widedeep.py:35-37   (B, 3072)
  probs            = Dense(2, softmax, 'simulator_reward')(obs)              widedeep.py:38
The env's obs_layer is whatever layer is NAMED simulator_obs (slate.py:232-237): here the 3072-wide concat.
"""
import numpy as np

OBS_DIM = 3072


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


class WideDeepOracle:
    """forward(seq int[B,2,64], dense f[B,432], cat int[B,21]) -> (obs [B,3072], probs [B,2])."""

    def __init__(self, weights, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.w = {k: np.asarray(v, dtype=self.dt) for k, v in weights.items()}      # no copy when the dtype already matches

    def forward(self, seq, dense, cat):
        w = self.w
        seq = np.asarray(seq).astype(np.int64)
        cat = np.asarray(cat).astype(np.int64)
        pooled = np.concatenate([w["emb_seq"][seq[:, i, :]].mean(axis=1) for i in range(seq.shape[1])], axis=-1).astype(self.dt)
        s = _elu(pooled @ w["fc_w"] + w["fc_b"]).astype(self.dt)
        x = _elu(np.asarray(dense).astype(self.dt) @ w["dense_w1"] + w["dense_b1"])
        x = _elu(x @ w["dense_w2"] + w["dense_b2"]).astype(self.dt)
        c = w["emb_cat"][cat].reshape(cat.shape[0], -1)
        obs = np.concatenate([s, x, c], axis=-1).astype(self.dt)
        z = obs @ w["rew_w"] + w["rew_b"]
        z = z - z.max(axis=-1, keepdims=True)
        p = np.exp(z)
        return obs, (p / p.sum(axis=-1, keepdims=True)).astype(self.dt)

    def obs_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[0].astype(np.float32)

    def reward_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[1].astype(np.float32)
