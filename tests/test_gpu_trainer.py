"""GPU: PPO / A2C iterations over the real CUDA env (device-resident rollouts)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _env(seq, B=64):
    from test_gpu_parity import _synthetic, make_env
    cfg, cat, log, w = _synthetic(B, seq, support_rllib_mask=True, is_eval=False, cache_size=4 * B)
    return make_env(cfg, seq, cat, log, w, output_format="torch")


@pytest.mark.parametrize("algo,seq", [("PPO", False), ("A2C", True)])
def test_trainer_runs_on_cuda_env(algo, seq):
    import torch
    from rl4rs_b200.trainer import get_rl_model
    env = _env(seq)
    tr = get_rl_model(algo, {}, env=env)
    r = [tr.train() for _ in range(2)]
    assert all(np.isfinite(x["total_loss"]) and np.isfinite(x["episode_reward_mean"]) for x in r)
    assert r[-1]["timesteps_total"] == 2 * 64 * env.config["max_steps"]
    # sampled actions respect the mask => most slates are valid => mean reward is positive
    assert r[-1]["episode_reward_mean"] > 0
    buf = tr.buf
    m = buf.mask.gather(2, buf.action.unsqueeze(-1)).squeeze(-1)
    assert bool((m == 1).all())
    assert tr.evaluate(1) >= 0
    assert torch.isfinite(tr.policy.flat).all()
