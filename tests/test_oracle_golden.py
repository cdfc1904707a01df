"""CPU: the oracle restatement (oracle/env_np.py + oracle/dien_np.py) against the fixtures the
reference's own code produced (tests/golden/make_golden.py).  Integer outputs bit-exact."""
import numpy as np
import pytest

from golden_util import Golden, golden_names, assert_close_rel
from oracle.dien_np import DienOracle
from oracle.env_np import OracleEnv


def roll_oracle(g, dtype=np.float32):
    cfg = g.config
    if "np_seed" in g.meta:
        np.random.seed(g.meta["np_seed"])
    env = OracleEnv(cfg, g.log, g.catalog, DienOracle(g.weights, dtype), seq=g.seq)
    return env


def check_oracle_against_record(env, cfg, a, n_episodes, name):
    """Replay the recorded actions through the oracle env and hold everything it exposes to the record made by the
    reference's own code: integers / features bit-exact, observations and rewards at the parity tolerance."""
    T = cfg["max_steps"]
    k = 0
    for ep in range(n_episodes):
        obs = env.reset()
        np.testing.assert_array_equal(np.asarray([int(u) for u in env.samples.user]), a["reset_user"][ep])
        for key, val in obs.items():
            ref = a["reset_" + key][ep]
            if key == "obs":
                assert_close_rel(val, ref, what="%s reset obs" % name)
            else:
                np.testing.assert_array_equal(val, ref, err_msg="reset " + key)
        for t in range(T):
            np.testing.assert_array_equal(env.offline_action, a["offline_action"][k])
            obs, reward, done, info = env.step(a["action_in"][k])
            seqs, dense, cat = env.samples.features()
            np.testing.assert_array_equal(seqs, a["seq"][k], err_msg="seq step %d" % k)
            np.testing.assert_array_equal(dense, a["dense"][k], err_msg="dense step %d" % k)
            np.testing.assert_array_equal(cat, a["cat"][k], err_msg="cat step %d" % k)
            assert dense.dtype == np.float32 and cat.dtype == np.int32 and seqs.dtype == np.int32
            np.testing.assert_array_equal(env.samples.prev_actions, a["prev_actions"][k])
            np.testing.assert_array_equal(env.samples.get_violation(), a["violation"][k])
            np.testing.assert_array_equal(done, a["done"][k])
            for key, val in obs.items():
                ref = a["step_" + key][k]
                if key == "obs":
                    assert_close_rel(val, ref, what="%s obs step %d" % (name, k))
                else:
                    np.testing.assert_array_equal(val, ref, err_msg="%s step %d" % (key, k))
            assert_close_rel(reward, a["reward"][k], what="%s reward step %d" % (name, k))
            assert_close_rel(env.offline_reward, a["offline_reward"][k], rtol=1e-12, what="offline_reward")
            if "click_p" in a and t == T - 1:
                assert_close_rel(np.stack([i["click_p"] for i in info]), a["click_p"][ep], what="click_p")
            k += 1


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(name):
    g = Golden(name)
    check_oracle_against_record(roll_oracle(g), g.config, g.arr, g.n_episodes, name)


def test_tutorial_known_answers():
    """what the reference's own code yields on the one real record."""
    g = Golden("tutorial_slate_rllib")
    known = g.meta["known"]
    assert g.arr["step_action_mask"].sum(-1)[:, 0].tolist() == known["mask_popcounts"]
    np.testing.assert_allclose(g.arr["offline_reward"][-1], [known["offline_reward"]] * 4, rtol=1e-12)
    assert g.arr["violation"][-1].tolist() == [1] * 4
    assert g.arr["cat"][-1, 0].tolist() == [64054, 50887, 66367, 44932, 59460, 20543, 83978, 50138,
                                            74820, 58670, 1, 3, 5, 29, 72, 53, 52, 164, 211, 172, 172]
    assert g.log.seq_len[0] == 111 and g.arr["seq"][0, 0, 0, :4].tolist() == [14, 139, 83, 83]
    assert g.arr["seq"][0, 0, 0, -3:].tolist() == [218, 236, 215]
    assert not g.arr["seq"][:, :, 1].any()


def test_f32_oracle_tracks_f64_oracle():
    """The f32 restatement (what the CUDA path is held to) against the f64 restatement."""
    g = Golden("slate_rllib_replay")
    e32, e64 = roll_oracle(g, np.float32), roll_oracle(g, np.float64)
    o32, o64 = e32.reset(), e64.reset()
    assert_close_rel(o32["obs"], o64["obs"], rtol=2e-5, what="reset")
    for t in range(g.config["max_steps"]):
        a = e32.offline_action
        o32, r32, _, _ = e32.step(a)
        o64, r64, _, _ = e64.step(a)
        assert_close_rel(o32["obs"], o64["obs"], rtol=2e-5, what="obs")
        assert_close_rel(r32, r64, rtol=2e-5, what="reward")
