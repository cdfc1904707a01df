"""gym 0.19 surface the reference env uses (base.py:178-217, rl4rs/__init__.py:5-18).

``gym`` is not installed in this image.  If it is importable we use it (and register the env
ids); otherwise a duck-typed stand-in provides Env / spaces.{Box,Discrete,Dict} / register / make
with the attributes RLlib-style callers read (shape, n, spaces, low, high, sample, contains).
"""
import numpy as np

try:  # pragma: no cover - not available in the build image
    import gym as _gym
    from gym import spaces
    Env = _gym.Env
    HAVE_GYM = True
except Exception:  # ImportError or a broken install
    _gym = None
    HAVE_GYM = False

    class Env(object):
        metadata = {}
        observation_space = None
        action_space = None

    class _Space(object):
        shape = ()

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

        def __repr__(self):
            return "Box(%s, %s, %s)" % (self.low, self.high, self.shape)

    class Discrete(_Space):
        def __init__(self, n):
            self.n = int(n)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return "Discrete(%d)" % self.n

    class Dict(_Space):
        def __init__(self, spaces=None, **kw):
            self.spaces = dict(spaces or {}, **kw)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def __getitem__(self, key):
            return self.spaces[key]

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __repr__(self):
            return "Dict(%s)" % ", ".join("%s:%r" % kv for kv in self.spaces.items())

    class spaces(object):  # noqa: N801 - mimics the gym.spaces module
        Box = Box
        Discrete = Discrete
        Dict = Dict

_REGISTRY = {}


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = entry_point
    if HAVE_GYM:  # pragma: no cover
        try:
            _gym.envs.registration.register(id=id, entry_point=entry_point, **kwargs)
        except Exception:
            pass


def make(id, **kwargs):
    """gym.make('SlateRecEnv-v0', recsim=sim) (README.md:12-13)."""
    entry = _REGISTRY[id]
    if isinstance(entry, str):
        mod, name = entry.split(":")
        entry = getattr(__import__(mod, fromlist=[name]), name)
    return entry(**kwargs)
