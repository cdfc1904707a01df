"""TEST INFRASTRUCTURE -- a CPU stand-in for rl4rs_b200.engine.Engine with the same attribute / method surface, backed
by the NumPy oracle (oracle/env_np.py + dien_np.py).  It lets the HOST layer of the product -- rl4rs_b200/env/{base,
slate,seqslate}.py, gymshim, the file-cursor sampler, the three output formats -- run end to end on a box without a GPU
and be held against the reference's own fixtures (tests/test_host_layer_cpu.py).  It is never importable from
rl4rs_b200/: the product has no CPU path (tests/test_capi_exports.py::test_no_cpu_fallback)."""
import numpy as np
import torch

from oracle.dien_np import DienOracle
from oracle.env_np import OracleState


class OracleEngine(object):
    def __init__(self, config, seq, catalog, weights, log, device=None):
        self.config, self.seq, self.catalog, self.log = config, bool(seq), catalog, log
        self.device = torch.device("cpu")
        self.B, self.T = int(config["batch_size"]), int(config["max_steps"])
        self.P, self.A = int(config.get("page_items", 9)), int(config["action_size"])
        g = lambda k: bool(config.get(k, False))
        self.conti, self.raw, self.rllib = g("support_conti_env"), g("rawstate_as_obs"), g("support_rllib_mask")
        self.d3rl = g("support_d3rl_mask") and not self.rllib
        self.info_fetch = g("simulator_info_fetch")
        self._oracle_cfg = dict(config)                   # the oracle state reads the caller's action_emb_size (slate.py:21-25)
        if g("support_onehot_action"):
            config["action_emb_size"] = self.A
        self.emb_dim = self.A if g("support_onehot_action") else int(config.get("action_emb_size", 32))
        self.action_emb = np.eye(self.A) if g("support_onehot_action") else catalog.action_emb(self.emb_dim)
        self.obs_dim = 256
        if config.get("algo", "dien") == "dnn":           # the env half is the same; the dnn network is much cheaper on a CPU
            from oracle.dnn_np import DnnOracle
            self.net = DnnOracle(weights, np.float32)
        else:
            self.net = DienOracle(weights, np.float32)
        self.st = None
        self.paid = False
        self.obs = self.mask = self.reward = self.cat = self.dense = self.seqf = self.click_p = self.masked = None

    # ---- what r4_reset / r4_step leave behind in the output buffers ---------------------------------------------
    def _publish(self):
        s = self.st
        seqs, dense, cat = s.features()
        if self.raw:
            self.cat, self.dense, self.seqf = torch.from_numpy(cat.copy()), torch.from_numpy(dense.copy()), torch.from_numpy(seqs.copy())
        else:
            self.obs = torch.from_numpy(self.net.obs_layer((seqs, dense, cat, None)))
        if self.rllib:
            self.mask = torch.from_numpy(s.full_mask().astype(np.uint8))
        if self.d3rl:
            if self.seq:                                                     # seqslate.py:18-23
                p0 = s.cur_steps // s.page_items * s.page_items
                pe = min(p0 + s.page_items - 1, s.max_steps - 1)
                ma = s.prev_actions[:, pe + 1 - s.page_items:pe + 1]
            else:
                ma = s.prev_actions
            self.masked = torch.from_numpy(ma.astype(np.int32))

    def reset(self, rows):
        rows = np.asarray(rows)
        if rows.shape != (self.B,):
            raise ValueError("reset needs %d row indices" % self.B)
        self.st = OracleState(self._oracle_cfg, self.log, self.catalog, rows, self.seq)
        self.reward = torch.zeros(self.B, dtype=torch.float64)
        self.paid = False
        self._publish()

    def step(self, action):
        s = self.st
        if s.cur_steps >= self.T:
            raise IndexError("step past max_steps")                          # slate.py:198
        a = action.numpy() if isinstance(action, torch.Tensor) else np.asarray(action)
        if not self.conti:
            a = a.reshape(-1)
            if a.shape != (self.B,):
                raise ValueError("discrete action must have %d entries" % self.B)
            if a.size and (a.min() < 0 or a.max() >= self.A):
                raise IndexError("action id out of range [0, %d)" % self.A)
        s.act(a)
        self._publish()
        c = s.cur_steps
        self.paid = (c % self.P == 0) if self.seq else (c >= self.T)
        reward = np.zeros(self.B)
        if self.paid:                                                        # slate.py:281-308 / seqslate.py:136-160
            if self.seq:
                zero = self.rllib or bool(self.config.get("support_d3rl_mask", False))
                price = s.get_price(s.prev_actions[:, :c])[:, -self.P:]
            else:
                zero, price = True, s.get_price(s.prev_actions)
            probs = self.net.reward_layer(s.complete_features() + (None,))[:, 1].reshape(self.B, -1)
            reward = np.sum(price * probs, axis=1)
            if zero:
                reward[s.get_violation() < 0.5] = 0
            if self.info_fetch:
                self.click_p = torch.from_numpy(probs.astype(np.float32))
        self.reward = torch.from_numpy(reward)

    @property
    def cur_steps(self):
        return int(self.st.cur_steps)

    def prev_actions(self):
        return torch.from_numpy(self.st.prev_actions.astype(np.int32))

    def offline_action(self):
        items = self.log.items[self.st.rows]
        a = items[:, self.st.cur_steps].astype(np.int64) if self.st.cur_steps < self.T else np.zeros(self.B, np.int64)
        return torch.from_numpy(a.astype(np.int32)), (torch.from_numpy(self.action_emb[a]) if self.conti else None)

    def offline_reward(self):
        return torch.from_numpy(np.asarray(self.st.offline_reward, dtype=np.float64))

    def violation(self):
        return torch.from_numpy(self.st.get_violation().astype(np.int32))

    def features(self):
        seqs, dense, cat = self.st.features()
        return torch.from_numpy(cat.copy()), torch.from_numpy(dense.copy()), torch.from_numpy(seqs.copy())

    def log_row_of(self, record):
        return list(self.log.lines).index(record)

    def to_host(self, **tensors):
        return {k: v.numpy().copy() for k, v in tensors.items()}

    def launch_count(self):
        return 0
