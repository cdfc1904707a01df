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


def test_rawstate_trainer_runs_on_cuda_env():
    """PPO_rawstate (modelfree_train.py:55-56,235-240): the env hands out the raw state (rawstate_as_obs + rllib mask) and the
    policy embeds it itself; the rollout stays on the device, the learner is the torch twin."""
    import torch
    from test_gpu_parity import _synthetic, make_env
    from rl4rs_b200.trainer import get_rl_model
    B = 64
    cfg, cat, log, w = _synthetic(B, False, support_rllib_mask=True, rawstate_as_obs=True, is_eval=False, cache_size=4 * B)
    env = make_env(cfg, False, cat, log, w, output_format="torch")
    tr = get_rl_model("PPO_rawstate", {"sgd_minibatch_size": 64, "num_sgd_iter": 2}, env=env)
    assert tr.rawstate and tr.buf.obs.shape == (cfg["max_steps"], B, 21 + 432 + 128)
    r = [tr.train() for _ in range(2)]
    assert all(np.isfinite(x["total_loss"]) and np.isfinite(x["episode_reward_mean"]) for x in r)
    assert r[-1]["episode_reward_mean"] > 0
    buf = tr.buf
    assert bool((buf.mask.gather(2, buf.action.unsqueeze(-1)) == 1).all())
    # the packed observation is the env's raw state: category ids of the first step = the log rows' user ids
    o = env.reset()
    packed = tr.policy.pack(o)
    assert torch.equal(packed[:, :21].long(), o["category_feature"].long()) and torch.equal(packed[:, 21 + 432:].long(), o["sequence_feature"].reshape(B, -1).long())
    assert torch.isfinite(tr.policy.flat).all()


def test_gae_kernel_matches_the_torch_loop_bitwise():
    import torch
    from rl4rs_b200.trainer import KernelOps, RolloutBuffer
    dev = torch.device("cuda")
    for T, B, gamma, lam in ((9, 1000, 1.0, 1.0), (27, 333, 0.99, 0.95), (1, 5, 0.5, 0.0)):
        buf = RolloutBuffer(T, B, 284, dev)
        g = torch.Generator().manual_seed(T)
        buf.reward.copy_(torch.randn(T, B, generator=g) * 10); buf.value.copy_(torch.randn(T, B, generator=g))
        t_ref, a_ref = buf.returns_and_advantages(gamma, lam)
        t_k, a_k = KernelOps(284, dev, 34973).gae(buf.reward, buf.value, gamma, lam)
        assert torch.equal(a_k, a_ref) and torch.equal(t_k, t_ref)


def _random_batch(n, A, dev, seed=0):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    obs = torch.randn(n, 256, generator=g).to(dev)
    mask = torch.zeros(n, A, dtype=torch.uint8)
    for i in range(n):
        lo, hi = [(1, 40), (40, 148), (148, A)][i % 3]
        mask[i, lo:hi] = 1
        mask[i, lo + (i % 5)] = 0
    return obs, mask.to(dev)


def test_policy_kernels_match_torch_autograd():
    """K12: r4_policy_act / r4_policy_grad / r4_adam_step against the torch implementation (autograd, torch Adam)."""
    import torch
    from rl4rs_b200.policy import MaskedPolicy
    from rl4rs_b200.trainer import KernelOps, PPOTrainer, A2CTrainer
    dev = torch.device("cuda")
    A, n = 284, 300
    pol = MaskedPolicy(A, dev, seed=3)
    with torch.no_grad():
        pol.flat.add_(0.05 * torch.randn_like(pol.flat))
    ops = KernelOps(A, dev, pol.n_params)
    obs, mask = _random_batch(n, A, dev)
    # ---- act: logits / value / logp, greedy action, sampling respects the mask and the distribution
    a = torch.empty(n, dtype=torch.int32, device=dev)
    lp, v, lg = torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, A, device=dev)
    ops.act(pol.flat, obs, mask, False, 1, a, lp, v, lg)
    ta, tlp, tv, tlg = pol.act(obs, mask, explore=False)
    assert torch.equal(a, ta)
    assert torch.allclose(v, tv, rtol=1e-4, atol=1e-5) and torch.allclose(lp, tlp, rtol=1e-4, atol=1e-5)
    ok = mask.bool()
    assert torch.allclose(lg[ok], tlg[ok], rtol=1e-4, atol=1e-5) and bool((lg[~ok] < -1e30).all())
    reps = 4000
    o1, m1 = obs[:1].repeat(reps, 1).contiguous(), mask[:1].repeat(reps, 1).contiguous()
    a1 = torch.empty(reps, dtype=torch.int32, device=dev)
    ops.act(pol.flat, o1, m1, True, 7, a1, torch.empty(reps, device=dev), torch.empty(reps, device=dev), None)
    assert bool(m1[0][a1.long()].all())
    p = torch.softmax(tlg[0], -1)
    freq = torch.bincount(a1.long(), minlength=A).float() / reps
    assert float((freq - p).abs().max()) < 4.5 * float((p.max() * (1 - p.max()) / reps) ** 0.5) + 2e-3
    # ---- gradients: PPO (mean) and A2C (sum) against autograd
    g = torch.Generator(device="cpu").manual_seed(5)
    act = torch.stack([torch.multinomial(mask[i].float().cpu(), 1, generator=g)[0] for i in range(n)]).to(dev)
    old_logits = (tlg + 0.3 * torch.randn(n, A, generator=g).to(dev) * ok).contiguous()
    old_logp = torch.log_softmax(old_logits, -1).gather(1, act.unsqueeze(1)).squeeze(1).contiguous()
    old_v = (tv + torch.randn(n, generator=g).to(dev)).contiguous()
    adv = torch.randn(n, generator=g).to(dev)
    tgt = (old_v + 600 * torch.randn(n, generator=g).to(dev)).contiguous()      # exercises the vf clip (500)

    class _E(object):
        config = {"max_steps": 3, "batch_size": 100, "action_size": A}
        sim = type("S", (), {"engine": type("X", (), {"device": dev})()})()

    for cls, mode in ((PPOTrainer, 0), (A2CTrainer, 1)):
        tr = cls({"entropy_coeff": 0.01, "use_kernels": False}, _E(), device=dev)
        tr.policy = pol
        if pol.flat.grad is not None:
            pol.flat.grad.zero_()
        if mode == 0:
            tr.kl_coeff = 0.2
            total, st = tr.loss(obs, mask, act, old_logp, old_logits, old_v, adv, tgt)
            hp = {"clip": 0.3, "vf_clip": 500.0, "vf_coeff": 0.5, "kl_coeff": 0.2, "ent_coeff": 0.01}
            inv_n, scale = 1.0 / n, 1.0 / n
        else:
            total, st = tr.loss(obs, mask, act, adv, tgt)
            hp = {"clip": 0.0, "vf_clip": 0.0, "vf_coeff": 0.5, "kl_coeff": 0.0, "ent_coeff": 0.01}
            inv_n, scale = 1.0, 1.0
        total.backward()
        ref = pol.flat.grad.detach().clone()
        ops.stats.zero_()
        ops.policy_grad(mode, pol.flat, (obs, mask, act, old_logp, old_logits, old_v, adv, tgt), None, 0, n, hp, inv_n, scale)
        err = (ops.grad - ref).abs().max() / ref.abs().max()
        assert float(err) < 2e-5, (cls.__name__, float(err))
        assert abs(float(ops.stats[4]) - float(total)) <= 1e-4 * abs(float(total)), (float(ops.stats[4]), float(total))
        # minibatch through an index list == the same rows gathered
        idx = torch.randperm(n, generator=g)[:128].to(dev)
        ops.policy_grad(mode, pol.flat, (obs, mask, act, old_logp, old_logits, old_v, adv, tgt), idx, 0, 128, hp,
                 1.0 / 128 if mode == 0 else 1.0, 1.0)
        g_idx = ops.grad.clone()
        sel = tuple(x[idx].contiguous() for x in (obs, mask, act, old_logp, old_logits, old_v, adv, tgt))
        ops.policy_grad(mode, pol.flat, sel, None, 0, 128, hp, 1.0 / 128 if mode == 0 else 1.0, 1.0)
        assert torch.equal(g_idx, ops.grad)                                   # deterministic reduction
        # multi-tile path (several tiles per CTA, shared-memory accumulator) against the one-tile-per-CTA path
        ops.policy_grad(mode, pol.flat, (obs, mask, act, old_logp, old_logits, old_v, adv, tgt), None, 0, n, hp, inv_n, scale)
        g_single = ops.grad.clone()
        ops.policy_grad(mode, pol.flat, (obs, mask, act, old_logp, old_logits, old_v, adv, tgt), None, 0, n, hp, inv_n, scale, G=5)
        assert float((ops.grad - g_single).abs().max() / g_single.abs().max()) < 1e-5
    # ---- r4_ppo_epoch (all minibatch steps in one call) == the per-step loop, bit for bit
    hp = {"clip": 0.3, "vf_clip": 500.0, "vf_coeff": 0.5, "kl_coeff": 0.2, "ent_coeff": 0.0}
    data = (obs, mask, act, old_logp, old_logits, old_v, adv, tgt)
    perm = torch.randperm(n, generator=g).to(dev)
    pa, pb = pol.flat.detach().clone(), pol.flat.detach().clone()
    oa, ob = KernelOps(A, dev, pol.n_params), KernelOps(A, dev, pol.n_params)
    assert oa.ppo_epoch(pa, data, perm, n, 64, hp, 1e-3, None) == n // 64 and oa.step == n // 64
    for s in range(0, n - 64 + 1, 64):
        ob.policy_grad(0, pb, data, perm, s, 64, hp, 1.0 / 64, 1.0 / 64)
        ob.adam(pb, 1e-3, 1.0, None)
    assert torch.equal(pa, pb) and torch.equal(oa.stats, ob.stats) and float((pa - pol.flat).abs().max()) > 1e-4
    # ---- Adam: three steps against torch.optim.Adam, with and without clipping
    for clip in (0.0, 0.5):
        p_t = pol.flat.detach().clone().requires_grad_(True)
        p_k = pol.flat.detach().clone()
        opt = torch.optim.Adam([p_t], lr=1e-3)
        ops2 = KernelOps(A, dev, pol.n_params)
        for it in range(3):
            gr = torch.randn(pol.n_params, generator=g).to(dev) * (it + 1)
            p_t.grad = gr.clone()
            if clip:
                torch.nn.utils.clip_grad_norm_([p_t], clip)
            opt.step()
            ops2.grad.copy_(gr)
            ops2.adam(p_k, 1e-3, 1.0, clip)
        assert torch.allclose(p_k, p_t.detach(), rtol=1e-5, atol=1e-7), float((p_k - p_t.detach()).abs().max())
