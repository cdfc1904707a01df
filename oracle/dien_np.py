"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's DIEN user-response simulator.

PARITY UNPINNED for the floating-point network: the reference holds no golden vector, test or
fixture for this arithmetic and its third-party halves are absent from
/root/reference:
  * deepctr==0.9.0 (environment.yml:147): ``DynamicGRU`` (-> TF1 ``GRUCell`` / deepctr
    ``VecAttGRUCell``) and ``AttentionSequencePoolingLayer`` (-> ``LocalActivationUnit``/``DNN``),
    call sites rl4rs/nets/utils.py:3,120-124;
  * tensorflow-gpu==1.15.0 (environment.yml:214): Embedding / Attention / Dense / ELU /
    GlobalAveragePooling1D, call sites rl4rs/nets/utils.py:20-25,50-53,113 and rl4rs/nets/dien.py:35-36.
The published algorithms of those layers are restated here and anchored on the reference's own
call sites.  The WIRING of the graph (below) is pinned: tests/test_reference_graph.py holds this oracle to
what the reference's own rl4rs/nets/dien.py + utils.py compute when run over the eager layer stand-ins of
oracle/tf_eager_stub.py (tests/golden/nets/reference_graph.npz); the arithmetic inside the third-party layers is not.  The integer/state-machine half of the path IS pinned (tests/golden, made by the
reference's own code through oracle/ref_harness.py).

Graph (rl4rs/nets/dien.py:8-45):
  category_feature = id_input_processing_attn(cat)        utils.py:16-25
  dense_feature    = dense_input_processing(dense)        utils.py:48-54
  sequence_feature = sequence_input_attn([seq, cat[:, -10:]])   utils.py:100-129
  obs   = Dense(256, ELU)(concat[sequence, dense, category])    dien.py:34-35  ('simulator_obs')
  probs = Dense(2, softmax)(obs)                                dien.py:36     ('simulator_reward')
"""
import numpy as np

SEQ_NUM = 2


def weight_shapes(cfg=None):
    """The W-table: TF1 variable shapes of the DIEN graph, by our flat names."""
    cfg = cfg or {}
    H = cfg.get("category_hash_size", 100000)
    E = cfg.get("emb_size", 128)
    U = cfg.get("hidden_units", 128)
    D = cfg.get("dense_feature_num", 432)
    C = cfg.get("category_feature_num", 21)
    shapes = {
        "emb_cat": (H, E), "emb_seq": (H, E),
        "dense_w1": (D, U), "dense_b1": (U,), "dense_w2": (U, U), "dense_b2": (U,),
        "obs_w": (2 * E * SEQ_NUM + U + E + C * E, 256), "obs_b": (256,),
        "rew_w": (256, 2), "rew_b": (2,),
    }
    for i in range(SEQ_NUM):
        shapes.update({
            "gru%d_wg" % i: (2 * E, 2 * E), "gru%d_bg" % i: (2 * E,),
            "gru%d_wc" % i: (2 * E, E), "gru%d_bc" % i: (E,),
            "att%d_w1" % i: (4 * E, 64), "att%d_b1" % i: (64,),
            "att%d_w2" % i: (64, 16), "att%d_b2" % i: (16,),
            "att%d_k" % i: (16, 1), "att%d_b" % i: (1,),
            "augru%d_wg" % i: (3 * E, 4 * E), "augru%d_bg" % i: (4 * E,),
            "augru%d_wc" % i: (3 * E, 2 * E), "augru%d_bc" % i: (2 * E,),
        })
    return shapes


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


class DienOracle:
    """forward(seq int[B,2,64], dense f[B,432], cat int[B,21]) -> (obs [B,256], probs [B,2])."""

    def __init__(self, weights, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.w = {k: np.asarray(v, dtype=self.dt) for k, v in weights.items()}      # no copy when the dtype already matches

    # -- utils.py:16-25 -------------------------------------------------------------------
    def category_feature(self, cat):
        emb = self.w["emb_cat"][cat]                               # (B,21,128)  Embedding
        s = np.matmul(emb, emb.transpose(0, 2, 1))                 # Attention(): no scale, no mask
        s = s - s.max(axis=-1, keepdims=True)
        p = np.exp(s)
        p = p / p.sum(axis=-1, keepdims=True)
        att = np.matmul(p, emb)                                    # (B,21,128)
        c1 = att.mean(axis=1)                                      # GlobalAveragePooling1D
        c2 = emb.reshape(emb.shape[0], -1)                         # Flatten
        return np.concatenate([c1, c2], axis=-1).astype(self.dt)

    # -- utils.py:48-54 (dropout inert at inference) ---------------------------------------
    def dense_feature(self, dense):
        w = self.w
        x = _elu(dense.astype(self.dt) @ w["dense_w1"] + w["dense_b1"])
        x = _elu(x @ w["dense_w2"] + w["dense_b2"])
        return x.astype(self.dt)

    # -- TF1 GRUCell via deepctr DynamicGRU (utils.py:120), h0 = 0, all maxlen steps --------
    def gru(self, i, x):
        w = self.w
        wg, bg, wc, bc = w["gru%d_wg" % i], w["gru%d_bg" % i], w["gru%d_wc" % i], w["gru%d_bc" % i]
        B, T, E = x.shape
        h = np.zeros((B, E), self.dt)
        out = np.empty((B, T, E), self.dt)
        for t in range(T):
            xt = x[:, t]
            g = _sigmoid(np.concatenate([xt, h], 1) @ wg + bg)
            r, u = g[:, :E], g[:, E:]
            c = np.tanh(np.concatenate([xt, r * h], 1) @ wc + bc)
            h = (u * h + (1 - u) * c).astype(self.dt)
            out[:, t] = h
        return out

    # -- deepctr AttentionSequencePoolingLayer(att_hidden_units=(64,16), return_score=True),
    #    weight_normalization=False default, mask all-true (utils.py:111,121-122) --------------
    def att_scores(self, i, q, keys):
        w = self.w
        B, T, E = keys.shape
        qq = np.broadcast_to(q[:, None, :], keys.shape)
        a = np.concatenate([qq, keys, qq - keys, qq * keys], axis=-1)          # (B,T,512)
        a = _sigmoid(a @ w["att%d_w1" % i] + w["att%d_b1" % i])                 # DNN layer 1, sigmoid
        a = _sigmoid(a @ w["att%d_w2" % i] + w["att%d_b2" % i])                 # DNN layer 2, sigmoid
        s = a @ w["att%d_k" % i] + w["att%d_b" % i]                             # linear 16 -> 1, raw
        return s[..., 0].astype(self.dt)                                       # (B,T)

    # -- deepctr VecAttGRUCell via DynamicGRU(gru_type='AUGRU') (utils.py:123-124) ------------
    def augru(self, i, x, scores):
        w = self.w
        wg, bg, wc, bc = (w["augru%d_wg" % i], w["augru%d_bg" % i],
                          w["augru%d_wc" % i], w["augru%d_bc" % i])
        B, T, E = x.shape
        U = wc.shape[1]
        h = np.zeros((B, U), self.dt)
        for t in range(T):
            xt = x[:, t]
            g = _sigmoid(np.concatenate([xt, h], 1) @ wg + bg)
            r, u = g[:, :U], g[:, U:]
            c = np.tanh(np.concatenate([xt, r * h], 1) @ wc + bc)
            u = (1.0 - scores[:, t:t + 1]) * u
            h = (u * h + (1 - u) * c).astype(self.dt)
        return h

    # -- utils.py:100-129 -------------------------------------------------------------------
    def sequence_feature(self, seq, slate_ids):
        es = self.w["emb_seq"]
        q = es[slate_ids].mean(axis=1).astype(self.dt)             # reduce_mean(axis=1) (B,128)
        outs = []
        for i in range(SEQ_NUM):
            x = es[seq[:, i, :]]                                   # (B,64,128)
            H = self.gru(i, x)
            s = self.att_scores(i, q, H)
            outs.append(self.augru(i, H, s))
        return np.concatenate(outs, axis=-1), q

    def features(self, seq, dense, cat):
        seq = np.asarray(seq).astype(np.int64)
        cat = np.asarray(cat).astype(np.int64)
        sf, _ = self.sequence_feature(seq, cat[:, -10:])            # dien.py:29-30,33
        df = self.dense_feature(np.asarray(dense))
        cf = self.category_feature(cat)
        return np.concatenate([sf, df, cf], axis=-1).astype(self.dt)   # dien.py:34

    def forward(self, seq, dense, cat):
        w = self.w
        allf = self.features(seq, dense, cat)
        obs = _elu(allf @ w["obs_w"] + w["obs_b"]).astype(self.dt)      # dien.py:35
        z = obs @ w["rew_w"] + w["rew_b"]                               # dien.py:36
        z = z - z.max(axis=-1, keepdims=True)
        p = np.exp(z)
        probs = (p / p.sum(axis=-1, keepdims=True)).astype(self.dt)
        return obs, probs

    # keras.backend.function stand-ins (slate.py:232-237): take feat = (seq, dense, cat, slate_label)
    def obs_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[0].astype(np.float32)

    def reward_layer(self, feat):
        return self.forward(feat[0], feat[1], feat[2])[1].astype(np.float32)
