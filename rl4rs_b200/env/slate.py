"""Host mirror of rl4rs/env/slate.py: SlateState (RecState plug-in) and SlateRecEnv (RecSimBase).

The state lives on the GPU (prev_actions, bit-packed action mask, special flag); this class is a
view that keeps the reference's attribute / method names.  ``act`` is the single device call of a
step; ``SlateRecEnv.obs_fn`` and ``forward`` format what that call produced.
"""
import numpy as np

from .base import RecSimBase, RecState
from ..synth import Catalog
from ..utils.datautil import FeatureUtil


class StateView(object):
    """What ``RecState.state`` hands to ``obs_fn`` (slate.py:90-106), materialised on demand from the device.

    The reference returns the raw 6-field state rows ``[role_id, [seq0, seq1], dense, category, slate_label, label]``
    (slate.py:67-83,203-213), wrapped in ``{'state', 'action_mask'}`` / ``{'state', 'masked_actions', 'cur_steps'}`` in
    the two mask modes.  The built-in ``SlateRecEnv.obs_fn`` never looks at it (the simulator pass of a step is fused
    behind ``act``), so nothing is copied unless a custom ``obs_fn`` plug-in reads it.  Rows arrive already padded the
    way ``FeatureUtil.feature_extraction`` pads them (datautil.py:43-65; padding is idempotent): dense f32[432],
    category i32[21], sequences i32[64] x 2."""

    def __init__(self, st):
        self._st, self._rows, self._cache = st, None, {}

    def rows(self):
        if self._rows is None:
            eng = self._st.engine
            cat, dense, seq = eng.features()
            h = eng.to_host(cat=cat, dense=dense, seq=seq)
            self._rows = [[0, [h["seq"][i, 0], h["seq"][i, 1]], h["dense"][i], h["cat"][i], [0] * 9, 0]
                          for i in range(eng.B)]
        return self._rows

    def keys(self):
        eng = self._st.engine
        if eng.rllib:
            return ["state", "action_mask"]
        if eng.d3rl:
            return ["state", "masked_actions", "cur_steps"]
        return []

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, k):
        st = self._st
        if isinstance(k, (int, np.integer, slice)):
            if self.keys():
                raise KeyError(k)
            return self.rows()[k]
        if k not in self.keys():
            raise KeyError(k)
        if k == "state":
            return self.rows()
        if k == "action_mask":
            return st.action_mask
        if k == "masked_actions":
            return st.prev_actions
        return np.full((st.batch_size, 1), st.cur_steps)

    def __len__(self):
        return len(self.keys()) or self._st.batch_size

    def __iter__(self):
        return iter(self.keys() or self.rows())


class SlateState(RecState):
    """slate.py:8-217.  Constructed like the reference's, ``state_cls(config, records)`` (base.py:68-71,92-100);
    ``records`` are the record strings when the log was ingested from text, else row indices into the resident log."""

    seq = False

    def __init__(self, config, records, engine=None):
        super().__init__(config, records)
        engine = engine if engine is not None else config.get("__engine__")
        if engine is None:
            raise ValueError("SlateState needs the simulator's engine (construct it through RecSimBase.sample)")
        self.engine = engine
        rows = config.get("__rows__")
        if rows is None or len(rows) != len(records):
            if len(records) and isinstance(records[0], str):         # a caller-made record list: resolve against the log
                rows = [engine.log_row_of(r) for r in records]
            else:
                rows = records
        self.rows = np.asarray(rows, dtype=np.int64)
        self.batch_size = config["batch_size"]
        self.action_size = config["action_size"]
        self.action_emb_size = engine.emb_dim
        self.max_steps = config["max_steps"]
        self.infos = [{} for _ in range(self.batch_size)] if engine.config.get("output_format", "list") == "list" else {}
        self.action_emb = engine.action_emb                           # slate.py:21-25
        self.special_items = engine.catalog.special_items             # slate.py:26
        self.location_mask = self._location_mask(self.action_size)
        engine.reset(self.rows)                                       # slate.py:16-19 on the device

    # ---- static helpers of the reference ------------------------------------------------------
    @staticmethod
    def _location_mask(action_size):
        m = np.zeros((4, action_size), dtype=np.int64)                # slate.py:60-64
        m[0, 1:40] = 1
        m[1, 40:148] = 1
        m[2, 148:] = 1
        m[3, 0] = 1
        return m

    @staticmethod
    def get_iteminfo_from_file(iteminfo_file, action_size, action_emb_size=32):
        """slate.py:28-53 -> (item_info_d, action_emb)."""
        cat = Catalog.from_file(iteminfo_file)
        d = {str(i): {"item_vec": cat.item_vec[i].tolist(), "price": float(cat.price[i]),
                      "location": int(cat.location[i])} for i in range(cat.action_size)}
        return d, cat.action_emb(action_emb_size)

    @staticmethod
    def get_mask_from_file(iteminfo_file, action_size):
        """slate.py:55-65 -> (location_mask, special_items)."""
        cat = Catalog.from_file(iteminfo_file)
        return SlateState._location_mask(action_size), [int(x) for x in cat.special_items]

    @staticmethod
    def get_nearest_neighbor(actions, action_emb, temperature=None):
        """slate.py:180-184 (host utility used by tutorial.ipynb:251-254; tiny, NumPy)."""
        return np.argmax(np.einsum("ij,kj->ik", np.asarray(actions, dtype=np.float64), action_emb), axis=1)

    @staticmethod
    def get_nearest_neighbor_with_mask(actions, action_emb, action_mask, temperature=None):
        """slate.py:186-191.  The env itself resolves actions on the GPU (k_act); this static
        mirror exists for callers that use it directly."""
        score = np.einsum("ij,kj->ik", np.asarray(actions, dtype=np.float64), action_emb)
        score[np.asarray(action_mask) < 0.5] = -2 ** 31
        return np.argmax(score, axis=1)

    # ---- device-backed attributes ----------------------------------------------------------------
    @property
    def cur_steps(self):
        return self.engine.cur_steps

    @property
    def prev_actions(self):
        return self.engine.prev_actions().cpu().numpy().astype(np.int64)

    @property
    def action_mask(self):
        """Combined mask of the current state (action_mask & location_mask & special_mask)."""
        if self.engine.mask is None:
            raise AttributeError("action masks are materialised only with support_rllib_mask")
        return self.engine.mask.cpu().numpy().astype(np.int64)

    @property
    def state(self):
        """slate.py:90-106, materialised lazily (see StateView)."""
        return StateView(self)

    @property
    def _state(self):
        return StateView(self).rows()

    @property
    def user(self):
        return [str(int(x)) for x in self.engine.log.session_id[self.rows]]   # slate.py:109-110

    @property
    def info(self):
        return self.infos

    def get_price(self, actions):
        return self.engine.catalog.price[np.asarray(actions)]          # slate.py:112-115

    def get_violation(self):
        return self.engine.violation().cpu().numpy().astype(np.int64)  # slate.py:133-147

    @property
    def offline_action(self):
        items, emb = self.engine.offline_action()                      # slate.py:149-162
        fmt = self.engine.config.get("output_format", "list")
        res = emb if self.engine.conti else items
        if fmt == "torch":
            return res
        res = self.engine.to_host(offline_action=res)["offline_action"]
        if fmt == "numpy":
            return res
        return [x for x in res] if self.engine.conti else res.tolist()

    @property
    def offline_reward(self):
        r = self.engine.offline_reward()                               # slate.py:164-174
        fmt = self.engine.config.get("output_format", "list")
        if fmt == "torch":
            return r
        r = self.engine.to_host(offline_reward=r)["offline_reward"]
        return r if fmt == "numpy" else r.tolist()

    def act(self, actions):
        """slate.py:193-214 -- plus, fused behind it, obs_fn's simulator pass and forward's reward."""
        self.engine.step(actions)

    def to_string(self):
        if len(self.records) and isinstance(self.records[0], str):
            return "\n".join(self.records)                                  # slate.py:216-217
        return "\n".join("log row %d (user %s)" % (r, u) for r, u in zip(self.rows, self.user))


class SlateRecEnv(RecSimBase):
    """slate.py:220-308.  ``config['model_file']`` is an .npz of the W-table, or
    ``config['weights']`` holds the arrays."""

    seq = False
    obs_dim = 256

    def __init__(self, config, state_cls=SlateState):
        self.max_steps = config["max_steps"]
        self.batch_size = config["batch_size"]
        self.FeatureUtil = FeatureUtil(config)               # slate.py:226 (plug-ins call self.FeatureUtil.feature_extraction)
        super().__init__(config, state_cls)
        self.obs_dim = self.engine.obs_dim                   # 256; 3072 for widedeep ('simulator_obs' is its concat layer)
        if config.get("support_d3rl_mask", False) and not config.get("support_rllib_mask", False) \
                and not config.get("rawstate_as_obs", False):
            self.obs_dim = self.engine.obs_dim + (self.engine.P if self.seq else self.max_steps) + 1   # slate.py:274-277

    def get_model(self, config):
        algo = config.get("algo", "dien")                              # slate.py:239-242: rl4rs/nets/<algo>.py
        if algo not in ("dien", "dnn", "widedeep", "lstm"):
            raise NotImplementedError("simulator %r is not a graph of rl4rs/nets: 'dien', 'dnn', 'widedeep' and 'lstm' are built "
                                      "(SURVEY.md section 8f n4)" % (algo,))
        w = config.get("weights")
        if w is None:
            w = self.load_model_file(config["model_file"], config)
        return w

    @staticmethod
    def load_model_file(model_file, config):
        """``model_file``: a TF1 ``tf.train.Saver`` prefix as the reference restores it (base.py:148-151; files
        ``<prefix>.index`` + ``<prefix>.data-*``, README.md:124-137), or an .npz of the W-table."""
        from ..utils import tf_checkpoint
        if tf_checkpoint.is_saver_prefix(model_file):
            load = {"dnn": tf_checkpoint.load_dnn_checkpoint, "widedeep": tf_checkpoint.load_widedeep_checkpoint,
                    "lstm": tf_checkpoint.load_lstm_checkpoint}.get(
                config.get("algo", "dien"), tf_checkpoint.load_dien_checkpoint)
            return load(model_file, config, name_map=config.get("variable_name_map"))
        return dict(np.load(model_file))

    # slate.py:244-279
    def obs_fn(self, state):
        eng = self.engine
        cfg = self.config
        fmt = self.output_format
        rllib = cfg.get("support_rllib_mask", False)
        d3rl = cfg.get("support_d3rl_mask", False) and not rllib
        if cfg.get("rawstate_as_obs", False):
            out = {"category_feature": eng.cat, "dense_feature": eng.dense, "sequence_feature": eng.seqf}
        else:
            out = {"obs": eng.obs}
        if rllib:
            out["action_mask"] = eng.mask
        if fmt == "torch":
            if d3rl and "obs" in out:
                import torch
                cs = torch.full((eng.B, 1), eng.cur_steps, dtype=torch.float64, device=eng.device)
                return torch.cat([out["obs"].double(), eng.masked.double(), cs], dim=-1)
            return {k: v.clone() for k, v in out.items()} if (rllib or "obs" not in out) else out["obs"].clone()
        if d3rl and "obs" in out:
            out["_masked"] = eng.masked
        host = eng.to_host(**out)
        if "action_mask" in host and fmt == "list":
            host["action_mask"] = host["action_mask"].astype(np.int64)   # reference dtype (np.int, slate.py:17)
        if d3rl and "obs" in host:
            ma = host.pop("_masked")
            cs = np.full((eng.B, 1), eng.cur_steps)
            return np.concatenate([host["obs"], ma, cs], axis=-1)
        if not rllib and "obs" in host:
            return host["obs"]
        if fmt == "numpy":
            return host
        keys = list(host.keys())
        if "action_mask" in keys:   # reference puts action_mask first (slate.py:258-261,270-273)
            keys = ["action_mask"] + [k for k in keys if k != "action_mask"]
        return [{k: host[k][i] for k in keys} for i in range(eng.B)]

    # slate.py:281-308
    def forward(self, model, samples):
        eng = self.engine
        fmt = self.output_format
        if eng.info_fetch and eng.paid:
            cp = eng.to_host(click_p=eng.click_p)["click_p"]
            if isinstance(samples.infos, list):
                for i in range(eng.B):
                    samples.infos[i].update({"click_p": cp[i]})
            else:
                samples.infos["click_p"] = cp
        if fmt == "torch":
            return eng.reward.clone()
        r = eng.to_host(reward=eng.reward)["reward"]
        return r.tolist() if fmt == "list" else r
