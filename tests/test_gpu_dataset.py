"""GPU: offline-dataset generation (batchrl_trainer.py:172-320) against the same loop over the CPU oracle."""
import numpy as np
import pytest

from golden_util import assert_close_rel

pytestmark = pytest.mark.gpu


def _oracle_dataset(cfg, seq, conti, log, cat, w, epochs):
    from oracle.dien_np import DienOracle
    from oracle.env_np import OracleEnv
    c = dict(cfg, support_d3rl_mask=True, support_rllib_mask=False)
    if conti:
        c["support_conti_env"] = True
    env = OracleEnv(c, log, cat, DienOracle(w, np.float32), seq=seq)
    B, T = c["batch_size"], c["max_steps"]
    adim = 32 if conti else 1
    O = np.zeros((epochs, B, T + 1, 266), np.float32); A = np.zeros((epochs, B, T + 1, adim), np.float32)
    R = np.zeros((epochs, B, T + 1), np.float32); D = np.zeros((epochs, B, T + 1), np.float32)
    for i in range(epochs):
        O[i, :, 0] = env.reset()["obs"]
        a = env.offline_action
        A[i, :, 0] = np.asarray(a).reshape(B, adim)
        for j in range(T):
            obs, _, done, _ = env.step(a)
            O[i, :, j + 1] = obs["obs"]
            a = env.offline_action
            A[i, :, j + 1] = np.asarray(a).reshape(B, adim)
            R[i, :, j + 1] = env.offline_reward
            D[i, :, j + 1] = done
    p = np.random.permutation(epochs)
    n = epochs * B * (T + 1)
    return O[p].reshape(n, -1), A[p].reshape(n, -1), R[p].reshape(n), D[p].reshape(n)


@pytest.mark.parametrize("seq,conti", [(False, False), (False, True), (True, False)])
def test_dataset_generation_matches_oracle(seq, conti, tmp_path):
    from test_gpu_parity import _synthetic
    from rl4rs_b200 import dataset
    B, epochs = 16, 3
    cfg, cat, log, w = _synthetic(B, seq)
    fn = {(False, False): dataset.data_generate_rl4rs_a, (False, True): dataset.data_generate_rl4rs_a_conti,
          (True, False): dataset.data_generate_rl4rs_b}[(seq, conti)]
    path = str(tmp_path / "ds.npz")
    np.random.seed(11)
    got = fn(dict(cfg, catalog=cat, log=log, weights=w), path, epochs=epochs)
    np.random.seed(11)
    O, A, R, D = _oracle_dataset(cfg, seq, conti, log, cat, w, epochs)
    T = cfg["max_steps"]
    assert got["observations"].shape == (epochs * B * (T + 1), 266) and got["observations"].dtype == np.float32
    assert_close_rel(got["observations"][:, :256], O[:, :256], what="dataset obs")
    np.testing.assert_array_equal(got["observations"][:, 256:], O[:, 256:])       # masked_actions | cur_steps
    np.testing.assert_array_equal(got["actions"], A)
    np.testing.assert_allclose(got["rewards"], R, rtol=1e-6)
    np.testing.assert_array_equal(got["terminals"], D)
    assert bool(got["discrete_action"]) == (not conti) and (got["terminals"].sum() == epochs * B)
    z = np.load(path)
    np.testing.assert_array_equal(z["actions"], got["actions"])
