"""The policy the reference trains on this env: RLlib's MyMaskActionsModel
(rl4rs/nets/rllib/rllib_mask_model.py:7-64): obs(256) -> FC 64 tanh -> 284 logits
+ max(log(action_mask), float32.min); the value head shares the 64-d hidden (vf_share_layers).

Exploration is RLlib SoftQ with temperature 1 = sample from softmax(logits)
(modelfree_train.py:398-402); evaluation uses argmax (explore=False, :412-414).
Parameters live in one flat f32 buffer so the data-parallel gradient all-reduce is ONE NCCL call
(34 973 parameters ~ 140 KB).
"""
import math

import torch

OBS, HID = 256, 64
FLOAT_MIN = torch.finfo(torch.float32).min


class MaskedPolicy(object):
    def __init__(self, action_size=284, device="cuda", seed=0):
        self.A = action_size
        self.device = torch.device(device)
        shapes = [("w1", (OBS, HID)), ("b1", (HID,)), ("w2", (HID, action_size)), ("b2", (action_size,)),
                  ("wv", (HID, 1)), ("bv", (1,))]
        n = sum(math.prod(s) for _, s in shapes)
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.device, requires_grad=True)
        self.views, off = {}, 0
        with torch.no_grad():
            for name, shape in shapes:
                k = math.prod(shape)
                v = self.flat[off:off + k].view(shape)
                if len(shape) == 2:       # RLlib normc_initializer(1.0) (0.01 for the output layers)
                    w = torch.randn(shape, generator=g)
                    std = 0.01 if name in ("w2", "wv") else 1.0
                    w = w * std / w.pow(2).sum(0, keepdim=True).sqrt()
                    v.copy_(w.to(self.device))
                off += k
        self._shapes, self.n_params = shapes, n

    def params(self):
        out, off = {}, 0
        for name, shape in self._shapes:
            k = math.prod(shape)
            out[name] = self.flat[off:off + k].view(shape)
            off += k
        return out

    def forward(self, obs, mask):
        """obs f32 [B,256], mask {0,1} [B,A] -> (masked logits [B,A], value [B])."""
        p = self.params()
        h = torch.tanh(obs @ p["w1"] + p["b1"])
        logits = h @ p["w2"] + p["b2"]
        inf_mask = torch.clamp(torch.log(mask.to(torch.float32)), min=FLOAT_MIN)   # rllib_mask_model.py:55-58
        value = (h @ p["wv"] + p["bv"]).squeeze(-1)
        return logits + inf_mask, value

    @torch.no_grad()
    def act(self, obs, mask, explore=True):
        """-> (action i32 [B], logp [B], value [B], masked logits [B,A])."""
        logits, value = self.forward(obs, mask)
        logp_all = torch.log_softmax(logits, dim=-1)
        if explore:
            a = torch.multinomial(logp_all.exp(), 1).squeeze(-1)
        else:
            a = logits.argmax(dim=-1)
        return a.to(torch.int32), logp_all.gather(1, a.long().unsqueeze(1)).squeeze(1), value, logits


class RawStatePolicy(object):
    """RLlib 'mask_model_rawstate' (rl4rs/nets/rllib/rllib_mask_model.py:67-115 over rllib_rawstate_model.py:25-86): the
    policy reads the RAW state -- category ids [21], dense features [432], sequence ids [2,64] (`rawstate_as_obs`,
    slate.py:246-253) -- through its own embedding tables:
        category = mean_t E_c[cat]            (utils.id_input_processing, nets/utils.py:7-14)
        dense    = ELU(ELU(x W1 + b1) W2 + b2) (nets/utils.py:48-54; the Dropout layers are inactive outside Keras fit)
        sequence = [mean_t E_s[seq_0] | mean_t E_s[seq_1]]   (nets/utils.py:56-77: ONE table for both sequences)
        context  = ELU([sequence | dense | category] Wc + bc)   (256)
        logits   = context Wo + bo + max(log(mask), float32.min) ;  value = context Wv + bv
    This is the functional twin for `*_rawstate` algorithms: plain torch ops + autograd (no hand-written kernel -- it is not
    on the benchmarked path; the 25.6 M-parameter embedding tables make its SGD step an HBM-bound dense Adam update).
    Observations travel through the trainer as ONE packed f32 row [cat 21 | dense 432 | seq 128] (ids < 2^24 are exact)."""

    def __init__(self, action_size=284, device="cuda", seed=0, config=None):
        cfg = config or {}
        self.A = action_size
        self.device = torch.device(device)
        self.H, self.E, self.U = cfg.get("category_hash_size", 100000), cfg.get("emb_size", 128), cfg.get("hidden_units", 128)
        self.C, self.D = cfg.get("category_feature_num", 21), cfg.get("dense_feature_num", 432)
        self.S, self.L = cfg.get("seq_num", 2), cfg.get("maxlen", 64)
        self.obs_dim = self.C + self.D + self.S * self.L
        E, U = self.E, self.U
        shapes = [("emb_cat", (self.H, E)), ("emb_seq", (self.H, E)), ("dw1", (self.D, U)), ("db1", (U,)), ("dw2", (U, U)),
                  ("db2", (U,)), ("wc", (self.S * E + U + E, 256)), ("bc", (256,)), ("w2", (256, action_size)),
                  ("b2", (action_size,)), ("wv", (256, 1)), ("bv", (1,))]
        n = sum(math.prod(s) for _, s in shapes)
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.device, requires_grad=True)
        off = 0
        with torch.no_grad():
            for name, shape in shapes:
                k = math.prod(shape)
                v = self.flat[off:off + k].view(shape)
                if name.startswith("emb_"):            # Keras Embedding: uniform(-0.05, 0.05)
                    v.copy_(((torch.rand(shape, generator=g) - 0.5) * 0.1).to(self.device))
                elif name in ("w2", "wv"):             # normc_initializer(0.01)
                    w = torch.randn(shape, generator=g)
                    v.copy_((w * 0.01 / w.pow(2).sum(0, keepdim=True).sqrt()).to(self.device))
                elif len(shape) == 2:                  # Keras Dense: glorot uniform
                    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                    v.copy_(((torch.rand(shape, generator=g) * 2 - 1) * lim).to(self.device))
                off += k
        self._shapes, self.n_params = shapes, n

    params = MaskedPolicy.params

    def pack(self, obs):
        """env observation dict (rawstate_as_obs, torch format) -> packed f32 [B, 581]."""
        B = obs["category_feature"].shape[0]
        return torch.cat([obs["category_feature"].reshape(B, -1).to(torch.float32), obs["dense_feature"].reshape(B, -1).to(torch.float32),
                          obs["sequence_feature"].reshape(B, -1).to(torch.float32)], dim=1)

    def forward(self, obs, mask):
        p = self.params()
        C, D = self.C, self.D
        cat = obs[:, :C].long()
        dense = obs[:, C:C + D]
        seq = obs[:, C + D:].long().view(-1, self.S, self.L)
        elu = torch.nn.functional.elu
        cfeat = torch.nn.functional.embedding(cat, p["emb_cat"]).mean(dim=1)
        x = elu(elu(dense @ p["dw1"] + p["db1"]) @ p["dw2"] + p["db2"])
        sfeat = torch.cat([torch.nn.functional.embedding(seq[:, i], p["emb_seq"]).mean(dim=1) for i in range(self.S)], dim=1)
        ctx = elu(torch.cat([sfeat, x, cfeat], dim=1) @ p["wc"] + p["bc"])
        logits = ctx @ p["w2"] + p["b2"]
        inf_mask = torch.clamp(torch.log(mask.to(torch.float32)), min=FLOAT_MIN)
        return logits + inf_mask, (ctx @ p["wv"] + p["bv"]).squeeze(-1)

    act = MaskedPolicy.act
