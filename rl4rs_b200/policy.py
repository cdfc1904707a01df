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
