"""Offline-RL dataset generation over the GPU-resident env ("next" row n2 of SURVEY.md section 8f).

Mirrors ``script/batchrl_trainer.py:172-320`` (``data_generate_rl4rs_a`` / ``_a_conti`` / ``_b`` / ``_b_conti``):
the env is replayed with the LOGGED policy (``env.offline_action``), and every episode contributes
``max_steps + 1`` entries -- the reset observation plus one per step -- of
  observation  f32 [256 + 9 + 1]   d3rl form: simulator obs | masked_actions | cur_steps (slate.py:274-277)
  action       the logged item id (or its embedding in the continuous variants)
  reward       ``env.offline_reward`` AFTER the step (0 for the reset entry)
  terminal     ``done`` of the step (0 for the reset entry)
Episodes are shuffled with ``np.random.permutation`` and flattened exactly like the reference.

What differs: the rollout and its buffers stay on the GPU (``output_format='torch'``); one device-to-host copy at
the end.  The result is returned as arrays and saved as ``.npz`` (keys observations, actions, rewards, terminals,
discrete_action), or -- for a ``*.h5`` path, when h5py is importable (it is not in this image) -- in the layout
``MDPDataset.dump`` writes.
"""
import numpy as np
import torch

from . import gymshim as gym
from .env.seqslate import SeqSlateRecEnv, SeqSlateState
from .env.slate import SlateRecEnv, SlateState


def _generate(config, seq, conti, datasetfile, epochs, total):
    cfg = dict(config)
    cfg["support_d3rl_mask"] = True
    cfg["support_rllib_mask"] = False                 # the d3rl observation form is the `elif` branch (slate.py:98)
    cfg["output_format"] = "torch"
    if conti:
        cfg["support_conti_env"] = 1
        if cfg.get("support_onehot_action", False):
            cfg["action_emb_size"] = cfg["action_size"]
    B = cfg["batch_size"]
    sim = (SeqSlateRecEnv(cfg, state_cls=SeqSlateState) if seq else SlateRecEnv(cfg, state_cls=SlateState))
    env = gym.make("SeqSlateRecEnv-v0" if seq else "SlateRecEnv-v0", recsim=sim)
    T = cfg["max_steps"]
    epoch = epochs if epochs is not None else total // B
    dev = sim.engine.device
    obs_dim = sim.obs_dim
    adim = sim.engine.emb_dim if conti else 1
    observations = torch.zeros((epoch, B, T + 1, obs_dim), dtype=torch.float32, device=dev)
    actions = torch.zeros((epoch, B, T + 1, adim), dtype=torch.float32, device=dev)
    rewards = torch.zeros((epoch, B, T + 1), dtype=torch.float32, device=dev)
    terminals = torch.zeros((epoch, B, T + 1), dtype=torch.float32, device=dev)
    for i in range(epoch):
        obs = env.reset()
        observations[i, :, 0] = obs
        action = env.offline_action
        actions[i, :, 0] = action.reshape(B, adim)
        for j in range(T):
            obs, reward, done, info = env.step(action)
            observations[i, :, j + 1] = obs
            action = env.offline_action
            actions[i, :, j + 1] = action.reshape(B, adim)
            rewards[i, :, j + 1] = env.offline_reward
            terminals[i, :, j + 1] = done
    p = torch.as_tensor(np.random.permutation(epoch), device=dev)
    n = epoch * B * (T + 1)
    out = {
        "observations": observations[p].reshape(n, -1).cpu().numpy(),
        "actions": actions[p].reshape(n, -1).cpu().numpy(),
        "rewards": rewards[p].reshape(n).cpu().numpy(),
        "terminals": terminals[p].reshape(n).cpu().numpy(),
        "discrete_action": np.asarray(not conti),
    }
    if datasetfile:
        save_dataset(datasetfile, out)
    return out


def save_dataset(datasetfile, out):
    """``*.h5`` -> the file ``d3rlpy.dataset.MDPDataset.dump`` writes (batchrl_trainer.py:214-217: datasets observations,
    actions, rewards, terminals, episode_terminals, discrete_action, version), so ``MDPDataset.load`` reads it back; needs
    h5py, which this image does not have.  Anything else -> ``.npz`` with the same arrays."""
    if str(datasetfile).endswith((".h5", ".hdf5")):
        try:
            import h5py
        except ImportError as exc:                      # pragma: no cover - h5py is absent offline
            raise ImportError("writing %s needs h5py; pass a .npz path to get the same arrays without it" % datasetfile) from exc
        discrete = bool(out["discrete_action"])
        acts = out["actions"].reshape(-1).astype(np.int32) if discrete else out["actions"]
        with h5py.File(datasetfile, "w") as f:        # pragma: no cover
            f.create_dataset("observations", data=out["observations"])
            f.create_dataset("actions", data=acts)
            f.create_dataset("rewards", data=out["rewards"])
            f.create_dataset("terminals", data=out["terminals"])
            f.create_dataset("episode_terminals", data=out["terminals"])
            f.create_dataset("discrete_action", data=discrete)
            f.create_dataset("version", data="1.0")
            f.flush()
        return datasetfile
    np.savez(datasetfile, **out)
    return datasetfile


def data_generate_rl4rs_a(config, datasetfile=None, epochs=None):
    """batchrl_trainer.py:172-217 (SlateRecEnv, discrete logged actions, 1 000 000 // batch episodes)."""
    return _generate(config, False, False, datasetfile, epochs, 1000000)


def data_generate_rl4rs_a_conti(config, datasetfile=None, epochs=None):
    """batchrl_trainer.py:220-270 (continuous actions = the logged items' embeddings)."""
    return _generate(config, False, True, datasetfile, epochs, 1000000)


def data_generate_rl4rs_b(config, datasetfile=None, epochs=None):
    """batchrl_trainer.py:272-320 (SeqSlateRecEnv, 500 000 // batch episodes of max_steps + 1 entries)."""
    return _generate(config, True, False, datasetfile, epochs, 500000)


def data_generate_rl4rs_b_conti(config, datasetfile=None, epochs=None):
    return _generate(config, True, True, datasetfile, epochs, 500000)
