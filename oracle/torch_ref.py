"""TEST INFRASTRUCTURE ONLY -- a SECOND, independently written restatement of the two simulator graphs (torch, f64).

oracle/dien_np.py / dnn_np.py are the oracle; with deepctr / TensorFlow absent there is no reference-held vector to pin
their floating-point half (parity unpinned).  This file re-derives the same graphs from the layer
definitions -- written against the formulas of SURVEY.md section 8a ("DIEN forward, exactly as wired") and the layer
sources named there, with torch primitives and a different decomposition (fused gate matrices split per gate, per-step
attention, einsum contractions) -- so that a transcription slip in either restatement shows up as a disagreement
(tests/test_oracle_cross.py).  It is NOT a pin: both could share a misunderstanding of deepctr 0.9.0.

Layer definitions followed (public sources, cited by call site):
  tf.keras.layers.Embedding / Attention(use_scale=False) / GlobalAveragePooling1D / Flatten      nets/utils.py:16-25
  tf.keras.layers.Dense + ELU(alpha=1)                                                           nets/utils.py:48-54
  tf.nn.rnn_cell.GRUCell (TF 1.15):  gate = sigmoid([x, h] Wg + bg) -> r, u = split(gate, 2);
      c = tanh([x, r*h] Wc + bc);  h' = u*h + (1-u)*c                                             nets/utils.py:120
  deepctr 0.9.0 LocalActivationUnit: DNN(sigmoid, (64,16))([q, k, q-k, q*k]) kernel + bias, raw score  nets/utils.py:121-122
  deepctr 0.9.0 VecAttGRUCell: as GRUCell, then u = (1 - att) * u;  h' = u*h + (1-u)*c            nets/utils.py:123-124
  heads: Dense(256, ELU, 'simulator_obs'), Dense(2, softmax, 'simulator_reward')                 nets/dien.py:34-36, dnn.py:34-36
"""
import torch
import torch.nn.functional as F


def _t(w, name):
    return torch.as_tensor(w[name], dtype=torch.float64)


def dnn_forward(w, dense, cat):
    cat = torch.as_tensor(cat, dtype=torch.long)
    pooled = F.embedding(cat, _t(w, "emb_cat")).sum(dim=1) / cat.shape[1]
    x = torch.as_tensor(dense, dtype=torch.float64)
    for i in (1, 2):
        x = F.elu(F.linear(x, _t(w, "dense_w%d" % i).T, _t(w, "dense_b%d" % i)))
    a = F.elu(F.linear(torch.cat([pooled, x], dim=1), _t(w, "fc_w").T, _t(w, "fc_b")))
    obs = F.elu(F.linear(a, _t(w, "obs_w").T, _t(w, "obs_b")))
    return obs.numpy(), F.softmax(F.linear(obs, _t(w, "rew_w").T, _t(w, "rew_b")), dim=-1).numpy()


def widedeep_forward(w, seq, dense, cat):
    seq = torch.as_tensor(seq, dtype=torch.long)
    cat = torch.as_tensor(cat, dtype=torch.long)
    es = _t(w, "emb_seq")
    pooled = torch.cat([F.embedding(seq[:, i], es).sum(dim=1) / seq.shape[2] for i in range(seq.shape[1])], dim=1)
    s = F.elu(F.linear(pooled, _t(w, "fc_w").T, _t(w, "fc_b")))
    x = torch.as_tensor(dense, dtype=torch.float64)
    for i in (1, 2):
        x = F.elu(F.linear(x, _t(w, "dense_w%d" % i).T, _t(w, "dense_b%d" % i)))
    obs = torch.cat([s, x, F.embedding(cat, _t(w, "emb_cat")).flatten(1)], dim=1)
    return obs.numpy(), F.softmax(F.linear(obs, _t(w, "rew_w").T, _t(w, "rew_b")), dim=-1).numpy()


def _keras_gru_last(x, k, rk, b):
    """Keras v1 GRU (hard sigmoid, reset_after False, gates [z | r | h]) written per gate from the layer definition."""
    U = rk.shape[0]
    h = torch.zeros(x.shape[0], U, dtype=torch.float64)
    hs = lambda v: torch.clamp(0.2 * v + 0.5, 0.0, 1.0)
    for t in range(x.shape[1]):
        xt = x[:, t]
        z = hs(xt @ k[:, :U] + b[:U] + h @ rk[:, :U])
        r = hs(xt @ k[:, U:2 * U] + b[U:2 * U] + h @ rk[:, U:2 * U])
        hh = torch.tanh(xt @ k[:, 2 * U:] + b[2 * U:] + (r * h) @ rk[:, 2 * U:])
        h = z * h + hh - z * hh
    return h


def lstm_forward(w, seq, dense, cat):
    seq = torch.as_tensor(seq, dtype=torch.long)
    cat = torch.as_tensor(cat, dtype=torch.long)
    ec = F.embedding(cat, _t(w, "emb_cat"))
    cg = _keras_gru_last(ec, _t(w, "cgru_k"), _t(w, "cgru_rk"), _t(w, "cgru_b"))
    x = torch.as_tensor(dense, dtype=torch.float64)
    for i in (1, 2):
        x = F.elu(F.linear(x, _t(w, "dense_w%d" % i).T, _t(w, "dense_b%d" % i)))
    es = _t(w, "emb_seq")
    sg = [_keras_gru_last(F.embedding(seq[:, i], es), _t(w, "sgru%d_k" % i), _t(w, "sgru%d_rk" % i), _t(w, "sgru%d_b" % i))
          for i in range(seq.shape[1])]
    allf = torch.cat(sg + [x, cg, ec.flatten(1)], dim=1)
    obs = F.elu(F.linear(allf, _t(w, "obs_w").T, _t(w, "obs_b")))
    return obs.numpy(), F.softmax(F.linear(obs, _t(w, "rew_w").T, _t(w, "rew_b")), dim=-1).numpy()


def _gru_step(x, h, wg, bg, wc, bc, att=None):
    """One step of TF1 GRUCell / deepctr VecAttGRUCell with the fused kernels split per gate and per input half."""
    nx, nh = x.shape[1], h.shape[1]
    wgx, wgh = wg[:nx], wg[nx:]                       # rows: [x ; h]
    wr_x, wu_x = wgx[:, :nh], wgx[:, nh:]             # columns: [r | u]
    wr_h, wu_h = wgh[:, :nh], wgh[:, nh:]
    r = torch.sigmoid(x @ wr_x + h @ wr_h + bg[:nh])
    u = torch.sigmoid(x @ wu_x + h @ wu_h + bg[nh:])
    c = torch.tanh(x @ wc[:nx] + (r * h) @ wc[nx:] + bc)
    if att is not None:
        u = u - att * u                               # (1 - att) * u
    return u * h + c - u * c                          # u*h + (1-u)*c


def dien_forward(w, seq, dense, cat):
    seq = torch.as_tensor(seq, dtype=torch.long)
    cat = torch.as_tensor(cat, dtype=torch.long)
    B = cat.shape[0]
    # category branch: softmax(E E^T) E, mean over the 21 positions, concatenated with the flattened embeddings
    e = F.embedding(cat, _t(w, "emb_cat"))
    att = torch.softmax(torch.einsum("bik,bjk->bij", e, e), dim=-1)
    cfeat = torch.cat([torch.einsum("bij,bjk->bik", att, e).mean(dim=1), e.reshape(B, -1)], dim=1)
    # dense tower
    x = torch.as_tensor(dense, dtype=torch.float64)
    for i in (1, 2):
        x = F.elu(x @ _t(w, "dense_w%d" % i) + _t(w, "dense_b%d" % i))
    # sequence branch
    es = _t(w, "emb_seq")
    q = F.embedding(cat[:, -10:], es).mean(dim=1)                      # slate ids = last 10 category slots (dien.py:29-30)
    finals = []
    for i in range(2):
        keys = F.embedding(seq[:, i, :], es)                          # (B, 64, 128)
        g = [_t(w, "gru%d_%s" % (i, k)) for k in ("wg", "bg", "wc", "bc")]
        a = [_t(w, "att%d_%s" % (i, k)) for k in ("w1", "b1", "w2", "b2", "k", "b")]
        u = [_t(w, "augru%d_%s" % (i, k)) for k in ("wg", "bg", "wc", "bc")]
        h1 = torch.zeros(B, 128, dtype=torch.float64)
        outs = []
        for t in range(keys.shape[1]):
            h1 = _gru_step(keys[:, t], h1, *g)
            outs.append(h1)
        h2 = torch.zeros(B, 256, dtype=torch.float64)
        for t, ht in enumerate(outs):
            z = torch.cat([q, ht, q - ht, q * ht], dim=1)
            z = torch.sigmoid(torch.sigmoid(z @ a[0] + a[1]) @ a[2] + a[3])
            score = z @ a[4] + a[5]                                     # (B, 1), raw: no softmax, weight_normalization=False
            h2 = _gru_step(ht, h2, *u, att=score)
        finals.append(h2)
    allf = torch.cat(finals + [x, cfeat], dim=1)                        # dien.py:34: [sequence | dense | category]
    obs = F.elu(allf @ _t(w, "obs_w") + _t(w, "obs_b"))
    return obs.numpy(), F.softmax(obs @ _t(w, "rew_w") + _t(w, "rew_b"), dim=-1).numpy()
