"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's batched env state machine.

CPU oracle for the SlateRecEnv / SeqSlateRecEnv ``reset -> step`` path.  It is a *restatement*
(vectorised NumPy over a structure-of-arrays log), not a copy: each method cites the reference
lines it follows.  It is pinned against the reference's OWN code (run through
``oracle/ref_harness.py``) by the fixtures in ``tests/golden/`` (``tests/test_oracle_golden.py``).
The floating-point simulator network comes from ``oracle/dien_np.py`` (parity unpinned, see there).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Because the reference's per-row CPython loops (slate.py:200-213,
117-131; datautil.py:40-51) are vectorised here, this port is *faster* than the reference on the
same cores: as a CPU baseline it is an upper bound on the reference's speed.
"""
import numpy as np

INT_MIN_FILL = -2 ** 31          # slate.py:189


class FileCursor(object):
    """Row-index emulation of RecDataBase's file pointer (base.py:75,82-108)."""

    def __init__(self, n_rows, cache_size):
        self.n = n_rows
        self.cache_size = cache_size
        self.pos = 0
        self.sample_list = []

    def reset(self, reset_file=False):
        # base.py:102-108
        self.sample_list = []
        if reset_file:
            self.pos = 0
        for _ in range(self.cache_size):
            # base.py:83-90: at EOF seek(0), discard one line, take the next
            if self.pos >= self.n:
                self.pos = 1
            self.sample_list.append(self.pos)
            self.pos += 1

    def sample(self, batch_size, is_eval):
        # base.py:92-98
        if is_eval:
            assert self.cache_size == batch_size
            assert len(self.sample_list) == batch_size
            return np.asarray(self.sample_list[:batch_size], dtype=np.int64)
        idx = np.random.randint(0, len(self.sample_list), batch_size)   # == np.random.choice(list, B)
        return np.asarray(self.sample_list, dtype=np.int64)[idx]


def location_mask_table(action_size):
    """slate.py:60-64."""
    m = np.zeros((4, action_size), dtype=np.int64)
    m[0, 1:40] = 1
    m[1, 40:148] = 1
    m[2, 148:] = 1
    m[3, 0] = 1
    return m


def nearest_neighbor_with_mask(actions, action_emb, action_mask):
    """slate.py:186-191 (f64 scores, fill -2**31, first-index argmax)."""
    score = np.einsum("ij,kj->ik", np.asarray(actions, dtype=np.float64), action_emb)
    score[action_mask < 0.5] = INT_MIN_FILL
    return np.argmax(score, axis=1)


def nearest_neighbor(actions, action_emb):
    """slate.py:180-184."""
    return np.argmax(np.einsum("ij,kj->ik", np.asarray(actions, dtype=np.float64), action_emb), axis=1)


class OracleState(object):
    """SlateState (slate.py:8-217) / SeqSlateState (seqslate.py:8-126) over SoA rows."""

    def __init__(self, config, log, catalog, rows, seq=False):
        self.config = config
        self.log, self.cat, self.rows = log, catalog, np.asarray(rows, dtype=np.int64)
        self.seq = seq
        self.B = config["batch_size"]
        self.A = config["action_size"]
        self.max_steps = config["max_steps"]
        self.page_items = config.get("page_items", 9)
        self.maxlen = config.get("maxlen", 64)
        self.n_dense = config.get("dense_feature_num", 432)
        self.n_cat = config.get("category_feature_num", 21)
        self.prev_actions = np.zeros((self.B, self.max_steps), dtype=np.int64)     # slate.py:16
        self.action_mask = np.ones((self.B, self.A), dtype=np.int64)               # slate.py:17
        self.special_mask = np.ones((self.B, self.A), dtype=np.int64)              # slate.py:18
        self.cur_steps = 0
        self.action_emb = catalog.action_emb(config.get("action_emb_size", 32))    # slate.py:21
        if config.get("support_onehot_action", False):                             # slate.py:22-25
            self.action_emb = np.eye(self.A)
        self.location_mask = location_mask_table(self.A)                           # slate.py:26
        self.special_items = catalog.special_items
        self.item_vec32 = catalog.item_vec.astype(np.float32)
        # initial state (slate.py:67-83) in feature form (datautil.py:34-69)
        self.user_dense = log.user_dense[self.rows]                                # f32 [B,32]
        self.user_cat = log.user_cat[self.rows].astype(np.int32)                   # i32 [B,10]
        self.seq0 = log.user_seq[self.rows].astype(np.int32)                       # pre-padded [B,64]
        self._cur_action = np.zeros(self.B, dtype=np.int64)
        self._has_acted = False
        self.infos = [{} for _ in range(self.B)]

    # ---- masks -----------------------------------------------------------------------------
    def _layer(self, steps):
        # slate.py:93 (cur_steps // 3) ; seqslate.py:16 (cur_steps % page_items // 3)
        return (steps % self.page_items // 3) if self.seq else (steps // 3)

    def full_mask(self):
        loc = self.location_mask[self._layer(self.cur_steps)][None, :]
        return self.action_mask & loc & self.special_mask

    # ---- act (slate.py:193-214, seqslate.py:92-126) ------------------------------------------
    def act(self, actions):
        if self.config.get("support_conti_env", False):
            actions = nearest_neighbor_with_mask(actions, self.action_emb, self.full_mask())
        actions = np.asarray(actions, dtype=np.int64).reshape(self.B)
        self.prev_actions[:, self.cur_steps] = actions
        self.action_mask[np.arange(self.B), actions] = 0
        has_special = np.isin(self.prev_actions, self.special_items).any(axis=1)   # whole history (Q8)
        self.special_mask[np.ix_(has_special, self.special_items)] = 0
        self._cur_action = actions
        self._act_step = self.cur_steps
        self._has_acted = True
        self.cur_steps += 1
        if self.seq and self.cur_steps % self.page_items == 0:                    # seqslate.py:124-126
            self.action_mask[:] = 1
            self.special_mask[:] = 1

    # ---- feature rows (act's state rebuild + datautil.py:34-69) -----------------------------
    def _window(self, step):
        """slate slots that go into dense/cat for the state built at ``step``."""
        if self.seq:
            p0 = step // self.page_items * self.page_items
            w = self.prev_actions[:, p0:p0 + self.page_items]
            sid = step // self.page_items + 1
            seq1 = self.prev_actions[:, :p0]
        else:
            w = self.prev_actions
            sid = 1                                            # slate.py:211 (Q12)
            seq1 = self.prev_actions[:, :0]
        return w, sid, seq1

    def _assemble(self, window, sid, seq1, cur_action):
        n = window.shape[0]
        dense = np.zeros((n, max(self.n_dense, 32 + 40 * (window.shape[1] + 1))), np.float32)
        dense[:, :32] = np.repeat(self.user_dense, n // self.B, axis=0) if n != self.B else self.user_dense
        k = window.shape[1]
        dense[:, 32:32 + 40 * k] = self.item_vec32[window].reshape(n, -1)
        dense[:, 32 + 40 * k:32 + 40 * (k + 1)] = self.item_vec32[cur_action]
        dense = dense[:, :self.n_dense]                         # post-truncate / post-pad
        ucat = np.repeat(self.user_cat, n // self.B, axis=0) if n != self.B else self.user_cat
        cat = np.concatenate([ucat, np.full((n, 1), sid, np.int64), window, cur_action[:, None]], axis=1)
        catp = np.zeros((n, max(self.n_cat, cat.shape[1])), np.int32)
        catp[:, :cat.shape[1]] = cat
        catp = catp[:, :self.n_cat]
        seqs = np.zeros((n, 2, self.maxlen), np.int32)
        seqs[:, 0] = np.repeat(self.seq0, n // self.B, axis=0) if n != self.B else self.seq0
        if seq1.shape[1] > 0:                                    # pre-pad, keep last maxlen
            s1 = seq1[:, -self.maxlen:]
            s1 = np.repeat(s1, n // self.B, axis=0) if n != self.B else s1
            seqs[:, 1, self.maxlen - s1.shape[1]:] = s1
        return seqs, dense, catp

    def features(self):
        """Current ``_state`` as (seq int32[B,2,64], dense f32[B,432], cat int32[B,21])."""
        if not self._has_acted:                                 # _init_state, slate.py:72-80
            seqs = np.zeros((self.B, 2, self.maxlen), np.int32)
            seqs[:, 0] = self.seq0
            dense = np.zeros((self.B, self.n_dense), np.float32)
            dense[:, :32] = self.user_dense
            cat = np.zeros((self.B, self.n_cat), np.int32)
            cat[:, :10] = self.user_cat
            return seqs, dense, cat
        w, sid, seq1 = self._window(self._act_step)
        return self._assemble(w, sid, seq1, self._cur_action)

    def complete_features(self):
        """Reward rows, (B*page) x features, env-row major (slate.py:117-131,289-292;
        seqslate.py:27-50,142-146)."""
        if self.seq:
            steps = range(self.cur_steps - self.page_items, self.cur_steps)
        else:
            steps = range(self.max_steps)
        outs = [self._assemble(*self._window(j), self.prev_actions[:, j]) for j in steps]
        P = len(outs)
        res = []
        for k in range(3):
            a = np.stack([o[k] for o in outs], axis=1)          # (B, P, ...)
            res.append(a.reshape((self.B * P,) + a.shape[2:]))
        return tuple(res)

    # ---- violation (slate.py:133-147, seqslate.py:52-69) ------------------------------------
    def get_violation(self):
        tmp = np.ones(self.B, dtype=np.int64)
        for step in range(self.cur_steps):
            tmp &= self.location_mask[self._layer(step)][self.prev_actions[:, step]]
        for step in range(max(self.cur_steps - 1, 1)):
            tmp &= (self.prev_actions[:, step] != self.prev_actions[:, step + 1])
        for step in range(max(self.cur_steps - 2, 1)):
            tmp &= (self.prev_actions[:, step] != self.prev_actions[:, step + 2])
        sp = np.zeros(self.A + 1, dtype=bool)
        sp[self.special_items] = True
        if self.seq:
            pages = range(self.cur_steps % self.page_items + 1)                    # Q9
            windows = [self.prev_actions[:, self.page_items * j:self.page_items * (j + 1)] for j in pages]
        else:
            windows = [self.prev_actions]
        for w in windows:
            for i in range(self.B):
                if len(np.intersect1d(w[i], self.special_items)) > 1:
                    tmp[i] = 0
        return tmp

    def get_price(self, actions):
        return self.cat.price[actions]                          # slate.py:112-115

    # ---- logged policy (slate.py:149-174, seqslate.py:71-86) --------------------------------
    @property
    def offline_action(self):
        items = self.log.items[self.rows]
        if self.cur_steps < self.max_steps:
            a = items[:, self.cur_steps].astype(np.int64)
        else:
            a = np.zeros(self.B, dtype=np.int64)
        if self.config.get("support_conti_env", False):
            return self.action_emb[a]
        return a

    @property
    def offline_reward(self):
        items = self.log.items[self.rows].astype(np.int64)
        fb = self.log.feedback[self.rows].astype(np.int64)
        c = self.cur_steps
        if self.seq:
            if c % 9 != 0 or c == 0:                            # c == 0: empty slices sum to 0
                return np.zeros(self.B)
            price = self.get_price(items[:, :c])[:, -self.page_items:]
            lab = fb[:, c - self.page_items:c]
            return np.sum(price * lab, axis=1)
        if c < self.max_steps:
            return np.zeros(self.B)
        return np.sum(self.get_price(items) * fb, axis=1)

    @property
    def user(self):
        return [str(int(x)) for x in self.log.session_id[self.rows]]   # slate.py:109-110 (Q13)


class OracleEnv(object):
    """RecEnvBase(SlateRecEnv|SeqSlateRecEnv) restated: base.py:157-175,181-269; slate.py:244-308;
    seqslate.py:136-160.  Observations are returned batched: a dict of arrays rather than the
    reference's list of per-row dicts (the host mirror in rl4rs_b200 does the list conversion)."""

    def __init__(self, config, log, catalog, dien, seq=False):
        self.config = dict(config)
        self.log, self.catalog, self.dien, self.seq = log, catalog, dien, seq
        self.B = config["batch_size"]
        self.max_steps = config["max_steps"]
        self.page_items = config.get("page_items", 9)
        self.cursor = FileCursor(log.n, config.get("cache_size", 2048))
        self.is_eval = config.get("is_eval", False)
        self.cur_step = 0
        # base.py:186-187 then :230 -- construction consumes two cache windows (Q19)
        self.cursor.reset()
        self._sample()
        self.reset()

    def seed(self, sd=0):
        np.random.seed(sd)                                       # base.py:78-80,153-155,232-234

    def _sample(self):
        rows = self.cursor.sample(self.B, self.is_eval)
        self.samples = OracleState(self.config, self.log, self.catalog, rows, self.seq)
        self.obs = self._obs()

    def reset(self, reset_file=False):
        self.cur_step = 0
        self.cursor.reset(reset_file)
        self._sample()
        return self.obs

    # slate.py:244-279
    def _obs(self):
        s = self.samples
        seqs, dense, cat = s.features()
        cfg = self.config
        out = {}
        if cfg.get("rawstate_as_obs", False):
            out.update(category_feature=cat, dense_feature=dense, sequence_feature=seqs)
        else:
            out["obs"] = self.dien.obs_layer((seqs, dense, cat, None))
        if cfg.get("support_rllib_mask", False):
            out["action_mask"] = s.full_mask()
        elif cfg.get("support_d3rl_mask", False) and not cfg.get("rawstate_as_obs", False):
            if self.seq:                                         # seqslate.py:18-23
                p0 = s.cur_steps // s.page_items * s.page_items
                pe = min(p0 + s.page_items - 1, s.max_steps - 1)
                ma = s.prev_actions[:, pe + 1 - s.page_items:pe + 1]
            else:
                ma = s.prev_actions
            cs = np.full((self.B, 1), s.cur_steps)
            out = {"obs": np.concatenate([out["obs"], ma, cs], axis=-1)}      # slate.py:274-277
        return out

    # slate.py:281-308 / seqslate.py:136-160
    def _reward(self):
        s = self.samples
        cfg = self.config
        if self.seq:
            if s.cur_steps % self.page_items != 0:
                return np.zeros(self.B), None
            zero_on_violation = cfg.get("support_rllib_mask", False) or cfg.get("support_d3rl_mask", False)
            price = s.get_price(s.prev_actions[:, :s.cur_steps])[:, -self.page_items:]
        else:
            if s.cur_steps < self.max_steps:
                return np.zeros(self.B), None
            zero_on_violation = True                              # slate.py:303 ``if 1:``
            price = s.get_price(s.prev_actions)
        feat = s.complete_features()
        probs = self.dien.reward_layer(feat + (None,))[:, 1].reshape(self.B, -1)
        reward = np.sum(price * probs, axis=1)
        if zero_on_violation:
            reward[s.get_violation() < 0.5] = 0
        return reward, probs

    def step(self, action):
        s = self.samples
        step = self.cur_step                                      # base.py:158 (Q1)
        s.act(action)
        self.obs = self._obs()
        reward, probs = self._reward()
        if probs is not None and self.config.get("simulator_info_fetch", False):
            for i in range(self.B):
                s.infos[i].update({"click_p": probs[i]})
        done = np.zeros(self.B, np.int64) if step < self.max_steps - 1 else np.ones(self.B, np.int64)
        self.cur_step += 1
        return self.obs, reward, done, s.infos

    @property
    def offline_action(self):
        return self.samples.offline_action

    @property
    def offline_reward(self):
        return self.samples.offline_reward
