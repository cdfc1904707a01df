"""CPU: PPO / A2C trainer logic on a fake env, checkpoint round trip, and the N>1 path
(world_size 2, gloo): sharded learners + ONE gradient all-reduce == single learner on the full batch."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl4rs_b200.policy import MaskedPolicy
from rl4rs_b200.trainer import A2CTrainer, PPOTrainer, RolloutBuffer, get_rl_model

A = 284


class FakeEnv(object):
    """Torch-format env protocol on CPU: random obs, layered masks, reward = f(actions) at the end."""

    def __init__(self, B, T=9, seed=0):
        self.config = {"max_steps": T, "batch_size": B, "action_size": A}
        self.B, self.T = B, T
        self.g = torch.Generator().manual_seed(seed)
        self.sim = type("S", (), {"engine": type("E", (), {"device": torch.device("cpu")})()})()

    def _obs(self):
        m = torch.zeros(self.B, A, dtype=torch.uint8)
        lo, hi = [(1, 40), (40, 148), (148, A)][min(self.t // 3, 2)]
        m[:, lo:hi] = 1
        return {"obs": torch.randn(self.B, 256, generator=self.g), "action_mask": m}

    def reset(self):
        self.t, self.acc = 0, torch.zeros(self.B, dtype=torch.float64)
        return self._obs()

    def step(self, a):
        self.acc += (a.to(torch.float64) % 7)
        self.t += 1
        done = self.t >= self.T
        r = self.acc.clone() if done else torch.zeros(self.B, dtype=torch.float64)
        return self._obs(), r, torch.full((self.B,), int(done)), {}


def test_policy_mask_and_sampling():
    pol = MaskedPolicy(A, "cpu", seed=1)
    assert pol.n_params == 256 * 64 + 64 + 64 * A + A + 64 + 1 == 34973
    env = FakeEnv(64)
    o = env.reset()
    a, logp, v, logits = pol.act(o["obs"], o["action_mask"])
    assert ((a >= 1) & (a < 40)).all() and (logp <= 0).all()
    g, _, _, _ = pol.act(o["obs"], o["action_mask"], explore=False)
    assert (g == logits.argmax(-1)).all()
    assert torch.isfinite(logits[:, 1:40]).all() and (logits[:, 40:] < -1e30).all()


def test_gae_lambda1_is_return_to_go_minus_value():
    buf = RolloutBuffer(4, 3, A, "cpu")
    buf.reward.copy_(torch.tensor([[0., 1, 0], [0, 0, 2], [0, 0, 0], [5, 0, 1]]))
    buf.value.copy_(torch.arange(12.).reshape(4, 3) * 0.1)
    ret, adv = buf.returns_and_advantages(1.0, 1.0)
    rtg = torch.flip(torch.cumsum(torch.flip(buf.reward, [0]), 0), [0])
    assert torch.allclose(ret, rtg) and torch.allclose(adv, rtg - buf.value)


@pytest.mark.parametrize("algo", ["PPO", "A2C"])
def test_training_improves_reward_and_checkpoint_roundtrip(algo):
    torch.manual_seed(0)
    env = FakeEnv(256, seed=3)
    tr = get_rl_model(algo, {"lr": 3e-3}, env=env, device="cpu")
    first = np.mean([tr.train()["episode_reward_mean"] for _ in range(3)])
    for _ in range(25):
        res = tr.train()
    assert res["timesteps_total"] == 28 * 9 * 256 and np.isfinite(res["total_loss"])
    assert tr.evaluate(2) > first + 2.0            # learns to pick ids with large id % 7
    d = tempfile.mkdtemp()
    path = tr.save(d)
    tr2 = get_rl_model(algo, {"lr": 3e-3}, env=FakeEnv(256, seed=3), device="cpu")
    tr2.restore(path)
    assert torch.equal(tr2.policy.flat, tr.policy.flat) and tr2.iteration == tr.iteration
    o = env.reset()
    np.testing.assert_array_equal(tr.compute_actions(o), tr2.compute_actions(o))
    rl = tr.compute_actions({i: {"obs": o["obs"][i].numpy(), "action_mask": o["action_mask"][i].numpy()} for i in range(4)})
    assert sorted(rl.keys()) == [0, 1, 2, 3]
    with pytest.raises(NotImplementedError):
        get_rl_model("DQN", {}, env=env)


def _fill(buf, seed, lo, hi):
    g = torch.Generator().manual_seed(seed)
    T, Bfull = buf.T, 32
    obs = torch.randn(T, Bfull, 256, generator=g)
    act = torch.randint(1, 40, (T, Bfull), generator=g)
    rew = torch.zeros(T, Bfull); rew[-1] = torch.rand(Bfull, generator=g) * 10
    mask = torch.zeros(T, Bfull, A, dtype=torch.uint8); mask[:, :, 1:40] = 1
    buf.obs.copy_(obs[:, lo:hi]); buf.action.copy_(act[:, lo:hi]); buf.reward.copy_(rew[:, lo:hi]); buf.mask.copy_(mask[:, lo:hi])


def _prepare(tr):
    with torch.no_grad():
        for t in range(tr.T):
            logits, v = tr.policy.forward(tr.buf.obs[t], tr.buf.mask[t])
            tr.buf.logits[t].copy_(logits); tr.buf.value[t].copy_(v)
            tr.buf.logp[t].copy_(torch.log_softmax(logits, -1).gather(1, tr.buf.action[t].unsqueeze(1)).squeeze(1))


def _worker(rank, world, algo, init_file, out_dir, mb=288):
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    per = 32 // world
    tr = get_rl_model(algo, {"sgd_minibatch_size": mb, "shuffle_sequences": False},      # the TOTAL over the ranks (RLlib)
                      env=FakeEnv(per), device="cpu")
    _fill(tr.buf, 7, rank * per, (rank + 1) * per)
    _prepare(tr)
    st = tr.learn(tr.buf)
    torch.save({"flat": tr.policy.flat.detach(), "grad": tr.policy.flat.grad.detach().clone(), "stats": st},
               os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("algo", ["A2C", "PPO"])
def test_world_size_2_gloo_matches_single_learner(algo):
    """Two ranks with half the env rows each + one gradient all-reduce per optimizer step must land
    on the same parameters as one learner holding all rows (A2C: one step on the whole batch; PPO:
    one minibatch = the whole batch so that the data order is identical)."""
    d = tempfile.mkdtemp()
    mp.spawn(_worker, args=(2, algo, os.path.join(d, "init"), d), nprocs=2, join=True)
    r0, r1 = torch.load(os.path.join(d, "r0.pt")), torch.load(os.path.join(d, "r1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])                      # replicas stay in sync
    single = get_rl_model(algo, {"sgd_minibatch_size": 288, "shuffle_sequences": False}, env=FakeEnv(32), device="cpu")
    _fill(single.buf, 7, 0, 32)
    _prepare(single)
    p0 = single.policy.flat.detach().clone()
    st = single.learn(single.buf)
    moved = (single.policy.flat.detach() - p0).abs().max()
    assert moved > 1e-5
    # the all-reduced gradient equals the single learner's gradient (fp32 summation order aside) ...
    g1, g2 = r0["grad"], single.policy.flat.grad.detach()
    assert torch.equal(r0["grad"], r1["grad"])
    assert (g1 - g2).abs().max() <= 1e-5 * g2.abs().max(), ((g1 - g2).abs().max(), g2.abs().max())
    # ... and so does the Adam update wherever the gradient is not at rounding-noise level
    # (Adam's first step is lr*sign(g): a coordinate whose gradient is ~0 may flip sign)
    big = g2.abs() > 1e-3 * g2.abs().max()
    assert torch.allclose(r0["flat"][big], single.policy.flat.detach()[big], atol=2e-6)
    assert abs(r0["stats"]["total_loss"] - st["total_loss"]) <= 1e-4 * max(1.0, abs(st["total_loss"]))


def test_world_size_2_minibatch_is_the_total_over_ranks():
    """RLlib semantics (modelfree_train.py:197: "Total SGD batch size across all devices"): with sgd_minibatch_size = 96
    two ranks take 48 of their own samples per SGD step, and three such steps land on the parameters of ONE learner
    that walks the full batch in minibatches of 96 (unshuffled, a minibatch = 3 time steps x all 32 rows = the union
    of the two ranks' minibatches of that step)."""
    d = tempfile.mkdtemp()
    mp.spawn(_worker, args=(2, "PPO", os.path.join(d, "init"), d, 96), nprocs=2, join=True)
    r0, r1 = torch.load(os.path.join(d, "r0.pt")), torch.load(os.path.join(d, "r1.pt"))
    assert torch.equal(r0["flat"], r1["flat"]) and r0["stats"]["sgd_steps"] == 3
    single = get_rl_model("PPO", {"sgd_minibatch_size": 96, "shuffle_sequences": False}, env=FakeEnv(32), device="cpu")
    _fill(single.buf, 7, 0, 32)
    _prepare(single)
    p0 = single.policy.flat.detach().clone()
    st = single.learn(single.buf)
    assert st["sgd_steps"] == 3
    step = (single.policy.flat.detach() - p0).abs()
    diff = (r0["flat"] - single.policy.flat.detach()).abs()
    # three Adam steps of lr 1e-4 move a coordinate by up to 3e-4; the two runs differ by fp32 summation order only,
    # except where a gradient coordinate is at rounding-noise level (Adam's sign-like first steps): bound the bulk
    assert step.max() > 1e-4
    assert (diff <= 2e-6).float().mean() > 0.99, float((diff <= 2e-6).float().mean())
    assert abs(r0["stats"]["total_loss"] - st["total_loss"]) <= 1e-4 * max(1.0, abs(st["total_loss"]))


class FakeRawEnv(FakeEnv):
    """rawstate_as_obs protocol (slate.py:246-253): the observation is the raw state dict + the action mask; the reward
    depends on a category id of the state, so only a policy that reads (and embeds) the raw ids can learn it."""

    HASH = 300

    def __init__(self, B, T=9, seed=0):
        super().__init__(B, T, seed)
        self.config.update({"rawstate_as_obs": True, "category_hash_size": self.HASH})
        self.uid = 7 + torch.randint(0, 2, (B,), generator=self.g)    # two kinds of users: ids 7 and 8

    def _obs(self):
        o = super()._obs()
        cat = self.uid[:, None].repeat(1, 21)             # every category slot carries the user id: an undiluted signal
        return {"category_feature": cat, "dense_feature": 0.1 * torch.randn(self.B, 432, generator=self.g),
                "sequence_feature": torch.zeros(self.B, 2, 64, dtype=torch.int64), "action_mask": o["action_mask"]}

    def step(self, a):
        # reward 1 per step whose action parity matches the user id's parity
        self.acc += ((a.to(torch.int64) % 2) == (self.uid % 2)).to(torch.float64)
        self.t += 1
        done = self.t >= self.T
        r = self.acc.clone() if done else torch.zeros(self.B, dtype=torch.float64)
        return self._obs(), r, torch.full((self.B,), int(done)), {}


def test_rawstate_policy_matches_layer_definitions_and_learns():
    from rl4rs_b200.policy import RawStatePolicy
    env = FakeRawEnv(96)
    pol = RawStatePolicy(A, "cpu", seed=2, config=env.config)
    assert pol.obs_dim == 21 + 432 + 128
    o = env.reset()
    packed = pol.pack(o)
    logits, value = pol.forward(packed, o["action_mask"])
    # the same graph written from the reference's layer list (rllib_rawstate_model.py:50-56, nets/utils.py:7-14,48-54,56-77)
    p = pol.params()
    elu = torch.nn.functional.elu
    c = p["emb_cat"][o["category_feature"]].mean(1)
    d = elu(elu(o["dense_feature"] @ p["dw1"] + p["db1"]) @ p["dw2"] + p["db2"])
    sq = torch.cat([p["emb_seq"][o["sequence_feature"][:, i]].mean(1) for i in range(2)], 1)
    ctx = elu(torch.cat([sq, d, c], 1) @ p["wc"] + p["bc"])
    ref = ctx @ p["w2"] + p["b2"]
    m = o["action_mask"].bool()
    assert torch.allclose(logits[m], ref[m], atol=1e-6) and (logits[~m] < -1e30).all()
    assert torch.allclose(value, (ctx @ p["wv"] + p["bv"]).squeeze(-1), atol=1e-6)
    tr = get_rl_model("PPO_rawstate", {"lr": 1e-3, "sgd_minibatch_size": 96, "num_sgd_iter": 6}, env=env, seed=3)
    assert tr.rawstate and not tr.use_kernels and tr.buf.obs.shape[-1] == 581
    first = tr.train()["episode_reward_mean"]
    for _ in range(24):
        last = tr.train()["episode_reward_mean"]
    assert last > first + 0.5, (first, last)        # chance level is 4.5 of 9
    acts = tr.compute_actions({k: v.numpy() for k, v in env.reset().items()})
    assert acts.shape == (96,)
