"""TEST INFRASTRUCTURE ONLY -- stub-import harness for the reference's own NumPy half.

Runs the *unmodified* reference classes (``rl4rs/env/base.py``, ``rl4rs/env/slate.py``,
``rl4rs/env/seqslate.py``, ``rl4rs/utils/datautil.py``) from ``/root/reference`` under
Python 3.12 / numpy 2.x by stubbing the modules this container lacks (gym, tensorflow)
-- the recipe of SURVEY.md Appendix C.  The TF session/graph half (``base.py:119-131``,
``slate.py:228-237``) is replaced by the NumPy DIEN restatement in ``oracle/dien_np.py``
plugged in as ``obs_layer`` / ``reward_layer`` so the reference's own
``SlateRecEnv.obs_fn`` / ``forward`` code runs (``slate.py:244-308``).

``/root/reference`` exists only in the build container, never on the GPU box: this module
is used exclusively by ``tests/golden/make_golden.py`` (fixture generation) and by the
``not gpu`` tests that re-check the fixtures when the reference is present.  Nothing in
the product (``rl4rs_b200/``) imports it.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RL4RS_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rl4rs", "env"))


def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
    """keras_preprocessing 1.1.2 ``pad_sequences`` semantics (call sites datautil.py:44,52,59)."""
    n = len(sequences)
    if maxlen is None:
        maxlen = max(len(s) for s in sequences)
    out = np.full((n, maxlen), value, dtype=dtype)
    for i, s in enumerate(sequences):
        s = list(s)
        if not len(s):
            continue
        trunc = s[-maxlen:] if truncating == "pre" else s[:maxlen]
        trunc = np.asarray(trunc, dtype=dtype)
        if padding == "post":
            out[i, : len(trunc)] = trunc
        else:
            out[i, -len(trunc):] = trunc
    return out


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class _Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


class _Dict:
    def __init__(self, spaces=None, **kw):
        self.spaces = dict(spaces or {}, **kw)


class _NumpyCompat(object):
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(obj, *a, **kw):
        try:
            return np.array(obj, *a, **kw)
        except ValueError:
            return np.array(obj, *a, dtype=object, **kw)


def install_stubs():
    """Install the sys.modules stubs (idempotent) and return the imported reference modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "int"):
        np.int = int  # reference uses the removed alias (slate.py:17,18,60,134)
    if "gym" not in sys.modules or getattr(sys.modules["gym"], "_r4_stub", False) is False:
        gym = types.ModuleType("gym")
        gym._r4_stub = True

        class Env(object):
            pass

        gym.Env = Env
        spaces = types.ModuleType("gym.spaces")
        spaces.Box, spaces.Discrete, spaces.Dict = _Box, _Discrete, _Dict
        gym.spaces = spaces
        gym.make = lambda env_id, **kw: sys.modules["rl4rs.env"].RecEnvBase(**kw)
        envs = types.ModuleType("gym.envs")
        reg = types.ModuleType("gym.envs.registration")
        reg.register = lambda **kw: None
        envs.registration = reg
        gym.envs = envs
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                            "gym.envs.registration": reg})
    if "tensorflow" not in sys.modules:
        tf = types.ModuleType("tensorflow")
        tf._r4_stub = True
        names = ["tensorflow.python", "tensorflow.python.data", "tensorflow.python.data.ops",
                 "tensorflow.keras", "tensorflow.keras.preprocessing",
                 "tensorflow.keras.preprocessing.sequence"]
        mods = {n: types.ModuleType(n) for n in names}
        mods["tensorflow.python.data.ops"].dataset_ops = types.SimpleNamespace()
        mods["tensorflow.keras.preprocessing.sequence"].pad_sequences = pad_sequences
        tf.python = mods["tensorflow.python"]
        tf.keras = mods["tensorflow.keras"]
        tf.keras.preprocessing = mods["tensorflow.keras.preprocessing"]
        tf.keras.preprocessing.sequence = mods["tensorflow.keras.preprocessing.sequence"]
        sys.modules["tensorflow"] = tf
        sys.modules.update(mods)
    # a stand-in without a __spec__ breaks importlib.util.find_spec(name) for every later caller (torch._dynamo probes
    # 'tensorflow' that way the first time an optimizer is built)
    import importlib.machinery
    for name, m in list(sys.modules.items()):
        if getattr(m, "_r4_stub", False) or name.startswith(("gym.", "tensorflow.")):
            if getattr(m, "__spec__", None) is None and isinstance(m, types.ModuleType):
                m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import rl4rs.env.base as ref_base  # noqa
    import rl4rs.env.slate as ref_slate  # noqa
    import rl4rs.env.seqslate as ref_seqslate  # noqa
    import rl4rs.utils.datautil as ref_datautil  # noqa
    # numpy >= 1.24 refuses to build ragged arrays implicitly; numpy 1.19 (environment.yml:83)
    # made an object array at slate.py:289 / seqslate.py:142.  Give the reference modules a numpy
    # proxy whose ``array`` restores that behaviour; every other attribute is numpy's own.
    for mod in (ref_slate, ref_seqslate):
        if not isinstance(mod.np, _NumpyCompat):
            mod.np = _NumpyCompat()
    return ref_base, ref_slate, ref_seqslate, ref_datautil


def make_reference_env(config, dien_fns, seq=False):
    """Build the reference's own RecEnvBase(SlateRecEnv|SeqSlateRecEnv) with the TF half replaced.

    ``dien_fns`` = (obs_layer, reward_layer): callables taking the 4-tuple ``feat`` produced by
    ``FeatureUtil.feature_extraction`` and returning f32[B,256] / f32[B,2] (slate.py:232-237).
    RecSimBase.__init__ (base.py:114-131) is bypassed exactly where it touches TF; everything
    else (RecDataBase, state classes, obs_fn, forward, RecEnvBase) is reference code.
    """
    ref_base, ref_slate, ref_seqslate, ref_datautil = install_stubs()
    sim_cls = ref_seqslate.SeqSlateRecEnv if seq else ref_slate.SlateRecEnv
    state_cls = ref_seqslate.SeqSlateState if seq else ref_slate.SlateState
    sim = sim_cls.__new__(sim_cls)
    # slate.py:223-227 / seqslate.py:132-134 / base.py:114-118,131 without the TF session
    sim.max_steps = config["max_steps"]
    sim.batch_size = config["batch_size"]
    sim.FeatureUtil = ref_datautil.FeatureUtil(config)
    sim.config = config
    sim.model = None
    sim._recData = ref_base.RecDataBase(config, state_cls)
    sim.page_items = config.get("page_items", 9)

    class _NullCtx:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    class _Sess:
        graph = types.SimpleNamespace(as_default=lambda: _NullCtx())

        def as_default(self):
            return _NullCtx()

    sim.sess = _Sess()
    sim.graph = _Sess.graph
    sim.obs_layer, sim.reward_layer = dien_fns
    env = ref_base.RecEnvBase(sim)
    return env
