"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's `lstm` user-response simulator
(config['algo'] = 'lstm', rl4rs/env/slate.py:239-242 -> rl4rs/nets/lstm.py:8-45).  Despite the name every recurrent
layer is a Keras GRU.

PARITY UNPINNED for the floating-point network (no reference-held vector; TensorFlow 1.15 absent).  The recurrent
layer is `tensorflow.keras.layers.GRU(units)` of TF 1.15 WITHOUT v2 behaviour (the reference is session-style TF1
code, rl4rs/env/base.py:113-151): keras/layers/recurrent.py GRU -- activation tanh, recurrent_activation
HARD sigmoid clip(0.2 x + 0.5, 0, 1), reset_after = False, one bias vector, gate order [z | r | h] in the fused
kernels, h0 = 0, no masking (Embedding has mask_zero = False), the LAST state is the output:
    z = hs(x Wz + h Uz + bz)    r = hs(x Wr + h Ur + br)    hh = tanh(x Wh + (r * h) Uh + bh)    h <- z h + (1 - z) hh

  category_feature = [GRU_0(Embedding_0(cat)) | Flatten(Embedding_0(cat))]        utils.py:28-36    (B, 128 + 21*128)
  dense_feature    = ELU(ELU(dense W1 + b1) W2 + b2)                              utils.py:48-54    (B, 128)
  sequence_feature = concat_i GRU_{1+i}(Embedding_1(seq[:, i]))                   utils.py:78-97    (B, 2*128)
                     ONE embedding layer shared by both sequences, one GRU layer per sequence
  obs = 'simulator_obs' = Dense(256, ELU)([sequence_feature, dense_feature, category_feature])   lstm.py:34-35
  probs            = Dense(2, softmax, 'simulator_reward')(obs)                   lstm.py:36
"""
import numpy as np

OBS_DIM = 256


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def _hard_sigmoid(x, dt):
    return np.clip(dt.type(0.2) * x + dt.type(0.5), 0, 1).astype(dt)


def keras_gru_last(x, k, rk, b, dt):
    """x [B,T,E] -> last state [B,U] of a Keras (TF 1.15, v1) GRU: kernel k [E,3U], recurrent kernel rk [U,3U], bias b [3U]."""
    B, T, _ = x.shape
    U = rk.shape[0]
    h = np.zeros((B, U), dt)
    xp = (x.reshape(B * T, -1) @ k + b).reshape(B, T, 3 * U).astype(dt)
    for t in range(T):
        g = (h @ rk[:, :2 * U]).astype(dt)
        z = _hard_sigmoid(xp[:, t, :U] + g[:, :U], dt)
        r = _hard_sigmoid(xp[:, t, U:2 * U] + g[:, U:], dt)
        hh = np.tanh(xp[:, t, 2 * U:] + (r * h) @ rk[:, 2 * U:]).astype(dt)
        h = (z * h + (1 - z) * hh).astype(dt)
    return h


class LstmOracle:
    """forward(seq int[B,2,64], dense f[B,432], cat int[B,21]) -> (obs [B,256], probs [B,2])."""

    def __init__(self, weights, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.w = {k: np.asarray(v, dtype=self.dt) for k, v in weights.items()}      # no copy when the dtype already matches

    def forward(self, seq, dense, cat):
        w, dt = self.w, self.dt
        seq = np.asarray(seq).astype(np.int64)
        cat = np.asarray(cat).astype(np.int64)
        ec = w["emb_cat"][cat]
        cg = keras_gru_last(ec, w["cgru_k"], w["cgru_rk"], w["cgru_b"], dt)
        x = _elu(np.asarray(dense).astype(dt) @ w["dense_w1"] + w["dense_b1"])
        x = _elu(x @ w["dense_w2"] + w["dense_b2"]).astype(dt)
        sg = [keras_gru_last(w["emb_seq"][seq[:, i, :]], w["sgru%d_k" % i], w["sgru%d_rk" % i], w["sgru%d_b" % i], dt)
              for i in range(seq.shape[1])]
        allf = np.concatenate(sg + [x, cg, ec.reshape(cat.shape[0], -1)], axis=-1).astype(dt)
        obs = _elu(allf @ w["obs_w"] + w["obs_b"]).astype(dt)
        z = obs @ w["rew_w"] + w["rew_b"]
        z = z - z.max(axis=-1, keepdims=True)
        p = np.exp(z)
        return obs, (p / p.sum(axis=-1, keepdims=True)).astype(dt)

    def obs_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[0].astype(np.float32)

    def reward_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[1].astype(np.float32)
