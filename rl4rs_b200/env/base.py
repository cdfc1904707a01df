"""Host mirror of rl4rs/env/base.py: RecState, RecDataBase, RecSimBase, RecEnvBase.

Same class names, constructor arguments, methods, properties, return conventions and error
behaviour as the reference, so callers written against it (README.md:8-21, simulator_eval.py,
batchrl_trainer.py:172-197, exact_k_train.py:75-87, rllib_vector_env.py) run unchanged.  What
differs is where the work happens: the file sampler is a cursor over row indices of an
HBM-resident log (no per-reset parsing), and ``act`` + ``obs_fn`` + ``forward`` of one step are
ONE call into the CUDA library (rl4rs_b200/engine.py) whose results the three methods hand out.

``config['output_format']`` (an addition; unknown keys are ignored by the reference):
  'list'  (default) reference-compatible Python lists / list of dicts;
  'numpy' batched host arrays (obs dict of arrays; action_mask stays uint8), no per-row Python objects;
  'torch' device tensors, no host synchronisation at all.
"""
from abc import ABC, abstractmethod

import numpy as np

from .. import gymshim as gym
from ..synth import Catalog, LogSoA
from ..utils.datautil import FeatureUtil


def single_elem_support(func):
    """batch_size == 1 returns scalars instead of 1-element lists (base.py:9-23).  The reference only ever sees lists;
    the 'numpy' / 'torch' output formats hand out arrays, tensors and dicts of them, so the unwrapping is per element:
    a 1-long list / tuple / array / tensor becomes its element, a dict is unwrapped value by value, everything else
    (the shared info dict, scalars) is passed through."""
    seq_types = (list, tuple, np.ndarray)

    def is_batch1(x):
        if isinstance(x, seq_types):
            return len(x) == 1
        return hasattr(x, "is_cuda") and x.dim() >= 1 and x.shape[0] == 1       # torch tensor

    def unwrap(x):
        if isinstance(x, dict):
            return {k: (v[0] if is_batch1(v) else v) for k, v in x.items()}
        return x[0] if is_batch1(x) else x

    def wrapper(*args, **kwargs):
        res = func(*args, **kwargs)
        if isinstance(res, tuple) and len(res) == 4:       # step(): (obs, reward, done, info), element by element
            if is_batch1(res[1]) or is_batch1(res[2]):     # (the reference hands back a LIST of 4 here: base.py:19-21)
                return [unwrap(x) for x in res]
            return res
        if isinstance(res, dict):                          # numpy / torch observation dict
            return unwrap(res) if all(is_batch1(v) for v in res.values()) else res
        if is_batch1(res):
            return res[0]
        if isinstance(res, seq_types) and len(res) and isinstance(res[0], seq_types) and len(res[0]) == 1:
            return [x[0] for x in res]
        return res

    return wrapper


class RecState(ABC):
    """State plug-in protocol (base.py:26-57)."""

    def __init__(self, config, records):
        self.config = config
        self.records = records

    @property
    @abstractmethod
    def state(self):
        pass

    @property
    @abstractmethod
    def user(self):
        pass

    @property
    @abstractmethod
    def info(self):
        pass

    @abstractmethod
    def act(self, actions):
        pass

    @abstractmethod
    def to_string(self):
        pass


class RecDataBase(object):
    """The file-based sampler of base.py:60-108 as a cursor over rows of the resident log.

    Semantics kept: ``cache_size`` windows read sequentially; at EOF the file pointer rewinds,
    one line is discarded and the next one is served (base.py:85-88); eval mode serves the first
    ``batch_size`` rows of the window and requires cache_size == batch_size (base.py:93-96);
    otherwise rows are drawn with replacement through the GLOBAL numpy RNG (base.py:98), which
    ``seed`` seeds (base.py:78-80)."""

    def __init__(self, config, state_cls, engine=None):
        self.config = config
        self.sample_list = []
        self.state_cls = state_cls
        self.is_eval = config.get("is_eval", False)
        self.cache_size = config.get("cache_size", 2048)
        self.engine = engine
        self.n_rows = engine.log.n
        self.pos = 0

    @staticmethod
    def seed(seed):
        np.random.seed(seed)

    def sample_cache(self, num):
        for _ in range(num):
            if self.pos >= self.n_rows:
                self.pos = 1 if self.n_rows > 1 else 0
            self.sample_list.append(self.pos)
            self.pos += 1

    def sample(self, batch_size):
        if self.is_eval:
            assert self.cache_size == batch_size
            assert len(self.sample_list) == batch_size
            rows = np.asarray(self.sample_list[:batch_size], dtype=np.int64)
        else:
            idx = np.random.randint(0, len(self.sample_list), batch_size)
            rows = np.asarray(self.sample_list, dtype=np.int64)[idx]
        # The reference hands the RECORD STRINGS to the state plug-in (base.py:92-100: state_cls(config, records)).  When
        # the log was ingested from text they are still available and are passed as such; the row indices travel beside
        # them so the device-backed state classes need no string look-up.  Logs that exist only as arrays (synthetic /
        # .npz) have no record text: `records` are then the row indices themselves.
        lines = getattr(self.engine.log, "lines", None)
        records = [lines[i] for i in rows] if lines is not None else rows
        self.config["__engine__"], self.config["__rows__"] = self.engine, rows
        try:
            return self.state_cls(self.config, records)
        finally:
            self.config.pop("__rows__", None)

    def reset(self, reset_file=False):
        self.sample_list = []
        if reset_file:
            self.pos = 0
        self.sample_cache(self.cache_size)


class RecSimBase(ABC):
    """Simulator plug-in (base.py:111-175).  ``get_model`` loads the DIEN W-table onto the GPU."""

    seq = False

    def __init__(self, config, state_cls):
        from ..engine import Engine   # needs CUDA; deferred so the host classes import on CPU
        self.config = config
        self.max_steps = config["max_steps"]
        self.batch_size = config["batch_size"]
        self.output_format = config.get("output_format", "list")
        if self.output_format not in ("list", "numpy", "torch"):
            raise ValueError("output_format must be 'list', 'numpy' or 'torch'")
        catalog = config.get("catalog")
        if catalog is None:
            catalog = Catalog.from_file(config["iteminfo_file"])
        log = config.get("log")
        if log is None:
            log = FeatureUtil.load_log(config["sample_file"], config.get("maxlen", 64))
        assert isinstance(log, LogSoA)
        self.model = self.get_model(config)
        self.engine = Engine(config, self.seq, catalog, self.model, log, device=config.get("device"))
        self._recData = RecDataBase(config, state_cls, self.engine)

    def reset(self, reset_file=False):
        self._recData.reset(reset_file)

    @abstractmethod
    def get_model(self, config):
        pass

    @abstractmethod
    def obs_fn(self, state):
        pass

    @abstractmethod
    def forward(self, model, samples):
        pass

    def reload_model(self, model_file):
        """base.py:148-151: load another checkpoint (Saver prefix or .npz W-table) into the running simulator."""
        w = self.load_model_file(model_file, self.config)
        self.engine._load_weights(w)
        self.model = w

    def seed(self, sd=0):
        self._recData.seed(sd)
        np.random.seed(sd)

    def _step(self, samples, action, **kwargs):
        step = kwargs["step"]
        samples.act(action)                       # one fused device call: act + obs + reward
        next_obs = self.obs_fn(samples.state)
        reward = self.forward(self.model, samples)
        next_info = samples.info
        done_val = 0 if step < self.max_steps - 1 else 1          # base.py:165-168 (pre-increment step)
        if self.output_format == "list":
            done = [done_val] * self.batch_size
        elif self.output_format == "numpy":
            done = np.full(self.batch_size, done_val, dtype=np.int64)
        else:
            import torch
            done = torch.full((self.batch_size,), done_val, dtype=torch.int64, device=self.engine.device)
        return next_obs, reward, done, next_info

    def sample(self, batch_size):
        samples = self._recData.sample(batch_size)
        obs = self.obs_fn(samples.state)
        return samples, obs


def build_spaces(config, obs_dim):
    """(observation_space, action_space) of an env, from the CONFIG alone.  The reference measures the first sampled
    observation (base.py:188-217); here the widths are configuration (the feature rows live in HBM and a 'torch' /
    'numpy' observation is one array per field, not a list of rows), the bounds and the key layout are the reference's:
    raw-state fields or one `obs` vector, wrapped with `action_mask` for RLlib; a Box of action_emb_size for the
    continuous env (base.py:214-215), Discrete(action_size) otherwise."""
    box, n_actions = gym.spaces.Box, config["action_size"]
    if config.get("rawstate_as_obs", False):
        fields = {"category_feature": (config.get("category_feature_num", 21),),
                  "dense_feature": (config.get("dense_feature_num", 432),),
                  "sequence_feature": (config.get("seq_num", 2), config.get("maxlen", 64))}
        parts = {name: box(-1000000.0, 1000000.0, shape=shape) for name, shape in fields.items()}
    else:
        parts = {"obs": box(-100000.0, 100000.0, shape=(obs_dim,))}
    if config.get("support_rllib_mask", False):
        observation = gym.spaces.Dict(dict({"action_mask": box(0, 1, shape=(n_actions,))}, **parts))
    else:
        observation = gym.spaces.Dict(parts) if "obs" not in parts else parts["obs"]
    if config.get("support_conti_env", False):
        action = box(-1, 1, shape=(config["action_emb_size"],))
    else:
        action = gym.spaces.Discrete(n_actions)
    return observation, action


class _PerEnvValue(object):
    """Read-only env attribute holding one value per env row; a batch of one hands out the bare element
    (the reference stacks @property on @single_elem_support four times, base.py:236-254)."""

    def __init__(self, source, doc):
        self._read = single_elem_support(source)
        self.__doc__ = doc

    def __get__(self, env, owner=None):
        return self if env is None else self._read(env)

    def __set__(self, env, value):
        raise AttributeError("read-only: %s" % self.__doc__)


class RecEnvBase(gym.Env):
    """gym env over a RecSimBase (base.py:178-273); registered as SlateRecEnv-v0 / SeqSlateRecEnv-v0.
    One env object = `batch_size` env rows stepping in lock-step on one GPU."""

    metadata = {"render.modes": ["human"]}

    state = _PerEnvValue(lambda env: env.obs, "observation of the last reset (base.py:236-239)")
    user_id = _PerEnvValue(lambda env: env.samples.user, "session ids of the sampled rows (base.py:241-244)")
    offline_action = _PerEnvValue(lambda env: env.samples.offline_action, "the logged action of this step (base.py:246-249)")
    offline_reward = _PerEnvValue(lambda env: env.samples.offline_reward, "the logged reward of this step (base.py:251-254)")

    def __init__(self, recsim):
        self.sim, self.config = recsim, recsim.config
        self.batch_size = self.config["batch_size"]
        self.observation_space, self.action_space = build_spaces(self.config, getattr(recsim, "obs_dim", 256))
        # the reference draws TWO batches while constructing (base.py:186-187 to size its spaces, then :230) and callers'
        # seeds / file cursors depend on that: keep both draws although the spaces no longer need the first
        self._draw(False)
        self.reset()

    def _draw(self, reset_file):
        self.cur_step = 0
        self.sim.reset(reset_file)
        self.samples, self.obs = self.sim.sample(self.batch_size)

    def seed(self, sd=0):
        self.sim.seed(sd)
        np.random.seed(sd)

    def reset(self, reset_file=False):
        self._draw(reset_file)
        return self.state

    @single_elem_support
    def step(self, action):
        batched = isinstance(action, (list, np.ndarray)) or hasattr(action, "is_cuda")      # device tensors pass through
        result = self.sim._step(self.samples, action if batched else [action], step=self.cur_step)
        self.cur_step += 1
        return tuple(result)                                      # (obs, reward, done, info)

    def render(self, mode="human", close=False):
        print("Current State:", "\n")
        print(self.samples.to_string())
