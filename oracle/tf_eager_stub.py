"""TEST INFRASTRUCTURE ONLY -- runs the REFERENCE'S OWN simulator graph code (rl4rs/nets/{dien,dnn,widedeep,lstm}.py and
rl4rs/nets/utils.py, imported from /root/reference, unmodified) eagerly over NumPy.

What this pins and what it does not.  The container has neither tensorflow 1.15 nor deepctr 0.9.0, so the reference's
`get_model(config)` cannot run as is.  Here the modules it imports (`tensorflow`, `tensorflow.keras.layers`,
`tensorflow.keras.models`, `deepctr.layers.sequence`) are replaced by eager stand-ins whose tensors are plain
ndarrays: `layers.Input(name=...)` hands out the fed array, every layer computes at call time, variables are looked up
in a TF1 checkpoint (a {variable name: array} dict) under the names Keras would give them.  So

  * the WIRING is the reference's: which input feeds which layer, `[:, -10:]`, the loop over sequences, the order of
    every `Concatenate`, which layer is called 'simulator_obs' / 'simulator_reward', and -- because Keras auto-names
    layers in construction order inside the fresh graph of base.py:119-121 -- the ORDER in which the variable scopes
    `embedding`, `dense_1`, `dynamic_gru_3`, `gru_2` ... come to exist.  That checks the restated graphs in
    oracle/{dien,dnn,widedeep,lstm}_np.py and the scope table of rl4rs_b200/utils/tf_checkpoint.py against the
    reference's real code (tests/test_reference_graph.py, tests/golden/nets/reference_graph.npz);
  * the LAYER ARITHMETIC below is still a restatement of the published third-party definitions (Keras 2.2.4-tf layers of
    TF 1.15; deepctr 0.9.0 `DynamicGRU`, `AttentionSequencePoolingLayer`, `LocalActivationUnit`, `DNN`,
    `VecAttGRUCell`; TF1 `GRUCell`), written a third time, layer by layer.  The simulator arithmetic therefore stays
    PARITY UNPINNED in the sense of SURVEY.md section 8c: no vector produced by TensorFlow itself is available.

`full_reference_stack(config)` goes one step further: inside it the reference's SlateRecEnv / SeqSlateRecEnv construct
themselves with their own __init__ (tf.Graph, get_model, tf.Session, tf.train.Saver().restore of a Saver-format checkpoint
read by rl4rs_b200.utils.tf_checkpoint, keras.backend.function), so a fixture can be replayed through the reference's
unmodified Python from RecEnvBase down to the graph definition (tests/test_reference_graph.py).

Nothing under rl4rs_b200/ imports this module; /root/reference exists only in the build container.
"""
import contextlib
import importlib
import importlib.machinery
import re
import sys
import types

import numpy as np

from oracle.ref_harness import REFERENCE_ROOT, reference_available

DT = np.float64


class _Ctx(object):
    def __init__(self, feed, checkpoint):
        self.feed, self.ckpt = feed, (dict(checkpoint) if checkpoint is not None else None)
        self.uids, self.layers, self.created, self.used, self.inputs, self.requested = {}, [], [], set(), [], {}


_CTX = None


def _snake(name):                               # keras.utils.generic_utils.to_snake_case
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


def _unique(base):                              # keras.backend.unique_object_name(..., zero_based=True), per graph
    k = _CTX.uids.get(base, 0)
    _CTX.uids[base] = k + 1
    return base if k == 0 else "%s_%d" % (base, k)


def _variable(scope, inner, shape):
    """The variable `scope/inner` of the checkpoint; when the inner name differs (deepctr's are best knowledge), the one
    variable under the top-level scope with that shape."""
    shape = tuple(int(s) for s in shape)
    top = scope.split("/")[0]
    full = scope + "/" + inner
    if _CTX.ckpt is not None and full not in _CTX.ckpt:
        hits = [k for k, v in _CTX.ckpt.items() if k.split("/")[0] == top and tuple(v.shape) == shape and k not in _CTX.used]
        if len(hits) != 1:
            raise KeyError("graph asks for %s %s: not in the checkpoint, %d same-shape candidates under %r %s (has: %s)"
                           % (full, shape, len(hits), top, hits, sorted(k for k in _CTX.ckpt if k.split("/")[0] == top)))
        full = hits[0]
    if _CTX.ckpt is None:                       # the build before Saver.restore (base.py:121): structure only
        _CTX.created.append((top, full, shape))
        return np.zeros(shape, DT)
    v = _CTX.ckpt[full]
    if tuple(v.shape) != shape:
        raise ValueError("%s has shape %s, the graph builds %s" % (full, v.shape, shape))
    _CTX.used.add(full)
    _CTX.created.append((top, full, shape))
    return np.asarray(v, DT)


def _sigmoid(x):
    with np.errstate(over="ignore"):            # exp(+big) -> inf -> 0.0, the value TF's sigmoid saturates to
        return 1.0 / (1.0 + np.exp(-x))


def _softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


class Layer(object):
    def __init__(self, name=None, **kwargs):
        self.name = name or _unique(_snake(type(self).__name__))
        self.output = None
        _CTX.layers.append(self)

    def weight(self, inner, shape, scope=None):
        return _variable(scope or self.name, inner, shape)

    def __call__(self, inputs, **kwargs):
        self.output = self.call(inputs)
        return self.output


# ---- tensorflow.keras.layers (Keras 2.2.4-tf as shipped in TF 1.15) ---------------------------------------------------
def Input(shape=None, dtype="float32", name=None, **kwargs):
    _CTX.inputs.append(name)
    if _CTX.feed is None:                       # structure-probing build: one dummy row
        x = np.zeros((1,) + tuple(shape))
    else:
        x = np.asarray(_CTX.feed[name])
    assert tuple(x.shape[1:]) == tuple(shape), (name, x.shape, shape)
    return x.astype(np.int64) if "int" in dtype else x.astype(DT)


class Lambda(Layer):
    def __init__(self, function, **kw):
        super().__init__(**kw)
        self.fn = function

    def call(self, x):
        return self.fn(x)


class ELU(Layer):
    def __init__(self, alpha=1.0, **kw):
        super().__init__(**kw)
        self.alpha = alpha

    def call(self, x):
        return np.where(x > 0, x, self.alpha * np.expm1(np.minimum(x, 0)))


class Embedding(Layer):                          # mask_zero = False: id 0 is an ordinary row
    def __init__(self, input_dim, output_dim, **kw):
        super().__init__(**kw)
        self.shape, self.table = (input_dim, output_dim), None

    def call(self, ids):
        if self.table is None:
            self.table = self.weight("embeddings", self.shape)
        return self.table[np.asarray(ids).astype(np.int64)]      # K.cast(inputs, 'int32') + gather


class Dense(Layer):
    def __init__(self, units, activation=None, **kw):
        super().__init__(**kw)
        self.units, self.act = units, activation

    def call(self, x):
        w, b = self.weight("kernel", (x.shape[-1], self.units)), self.weight("bias", (self.units,))
        y = x @ w + b
        if self.act is None:
            return y
        return _softmax(y) if self.act == "softmax" else self.act(y)


class Dropout(Layer):                            # inference: identity
    def __init__(self, rate, **kw):
        super().__init__(**kw)

    def call(self, x):
        return x


class GlobalAveragePooling1D(Layer):             # no mask reaches it (mask_zero = False)
    def call(self, x):
        return x.mean(axis=1)


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, xs):
        return np.concatenate(list(xs), axis=self.axis)


class Permute(Layer):
    def __init__(self, dims, **kw):
        super().__init__(**kw)
        self.dims = tuple(dims)

    def call(self, x):
        return np.transpose(x, (0,) + self.dims)


class Attention(Layer):                          # dot-product attention, use_scale = False, no mask, [query, value] (key = value)
    def call(self, qv):
        q, v = qv
        return _softmax(q @ np.swapaxes(v, 1, 2)) @ v


class GRU(Layer):
    """keras.layers.GRU of TF 1.15 without v2 behaviour (recurrent.GRU): tanh, HARD sigmoid, reset_after = False,
    implementation 1, gate order z | r | h, h0 = 0, return_sequences = False."""

    def __init__(self, units, **kw):
        super().__init__(**kw)
        self.units = units

    def call(self, x):
        u = self.units
        k, rk, b = self.weight("kernel", (x.shape[-1], 3 * u)), self.weight("recurrent_kernel", (u, 3 * u)), self.weight("bias", (3 * u,))
        hs = lambda a: np.clip(0.2 * a + 0.5, 0.0, 1.0)
        h = np.zeros((x.shape[0], u), DT)
        for t in range(x.shape[1]):
            xz, xr, xh = np.split(x[:, t] @ k + b, 3, axis=1)
            z = hs(xz + h @ rk[:, :u])
            r = hs(xr + h @ rk[:, u:2 * u])
            hh = np.tanh(xh + (r * h) @ rk[:, 2 * u:])
            h = z * h + (1 - z) * hh
        return h


# ---- deepctr 0.9.0 layers.sequence / layers.core / contrib.rnn_v2 + utils ---------------------------------------------
class DNN(Layer):
    def __init__(self, hidden_units, activation, parent):
        super().__init__()
        self.hidden, self.act, self.parent = hidden_units, activation, parent

    def call(self, x):
        assert self.act == "sigmoid"
        for i, n in enumerate(self.hidden):
            w = self.weight("local_activation_unit/dnn/kernel%d" % i, (x.shape[-1], n), scope=self.parent)
            b = self.weight("local_activation_unit/dnn/bias%d" % i, (n,), scope=self.parent)
            x = _sigmoid(np.tensordot(x, w, axes=(-1, 0)) + b)            # use_bn False, dropout 0
        return x


class LocalActivationUnit(Layer):
    def __init__(self, hidden_units, activation, parent):
        super().__init__()
        self.hidden, self.parent = hidden_units, parent
        self.dnn = DNN(hidden_units, activation, parent)

    def call(self, inputs):
        query, keys = inputs                                               # (B,1,E), (B,T,E)
        q = np.repeat(query, keys.shape[1], axis=1)                        # K.repeat_elements
        a = self.dnn(np.concatenate([q, keys, q - keys, q * keys], axis=-1))
        k = self.weight("local_activation_unit/kernel", (self.hidden[-1], 1), scope=self.parent)
        b = self.weight("local_activation_unit/bias", (1,), scope=self.parent)
        return np.tensordot(a, k, axes=(-1, 0)) + b                        # (B,T,1)


class AttentionSequencePoolingLayer(Layer):
    def __init__(self, att_hidden_units=(80, 40), att_activation="sigmoid", weight_normalization=False,
                 return_score=False, supports_masking=False, **kw):
        super().__init__(**kw)
        self.norm, self.ret_score = weight_normalization, return_score
        self.lau = LocalActivationUnit(tuple(att_hidden_units), att_activation, self.name)

    def call(self, inputs):
        queries, keys, keys_length = inputs
        T = keys.shape[1]
        mask = (np.arange(T)[None, None, :] < np.asarray(keys_length).reshape(-1, 1, 1))      # tf.sequence_mask -> (B,1,T)
        out = np.transpose(self.lau([queries, keys]), (0, 2, 1))           # (B,1,T)
        out = np.where(mask, out, (-2.0 ** 32 + 1) if self.norm else 0.0)
        if self.norm:
            out = _softmax(out)
        return out if self.ret_score else out @ keys


class DynamicGRU(Layer):
    """dynamic_rnn over a TF1 GRUCell ('GRU') or deepctr's VecAttGRUCell ('AUGRU'): gates = sigmoid([x, h] Wg + bg),
    r, u = split(gates); c = tanh([x, r * h] Wc + bc); AUGRU: u <- (1 - att) u; h <- u h + (1 - u) c.  Steps at or
    beyond a row's sequence_length copy the state through and emit zeros (dynamic_rnn)."""

    def __init__(self, num_units=None, gru_type="GRU", return_sequence=True, **kw):
        super().__init__(**kw)
        assert gru_type in ("GRU", "AUGRU")
        self.units, self.kind, self.ret_seq = num_units, gru_type, return_sequence

    def call(self, inputs):
        x, length = inputs[0], np.asarray(inputs[1]).reshape(-1)
        att = inputs[2] if self.kind == "AUGRU" else None                  # (B,T,1)
        n = self.units or x.shape[-1]
        cell = "gru_cell" if self.kind == "GRU" else "vec_att_gru_cell"
        wg, bg = self.weight(cell + "/gates/kernel", (x.shape[-1] + n, 2 * n)), self.weight(cell + "/gates/bias", (2 * n,))
        wc, bc = self.weight(cell + "/candidate/kernel", (x.shape[-1] + n, n)), self.weight(cell + "/candidate/bias", (n,))
        h = np.zeros((x.shape[0], n), DT)
        outs = np.zeros((x.shape[0], x.shape[1], n), DT)
        for t in range(x.shape[1]):
            g = _sigmoid(np.concatenate([x[:, t], h], axis=1) @ wg + bg)
            r, u = g[:, :n], g[:, n:]
            c = np.tanh(np.concatenate([x[:, t], r * h], axis=1) @ wc + bc)
            if att is not None:
                u = (1.0 - att[:, t]) * u
            new_h = u * h + (1 - u) * c
            live = (t < length)[:, None]
            h = np.where(live, new_h, h)
            outs[:, t] = np.where(live, new_h, 0.0)
        return outs if self.ret_seq else h[:, None, :]                     # tf.expand_dims(hidden_state, axis=1)


class Model(object):
    def __init__(self, inputs=None, outputs=None, **kw):
        self.input, self.outputs = inputs, outputs
        self.layers = list(_CTX.layers)

    def get_layer(self, name):
        layer = [l for l in self.layers if l.name == name][0]
        if _CTX is not None:
            _CTX.requested[id(layer.output)] = name          # an activation layer shares its output array with its Dense
        return layer

    def compile(self, **kw):
        pass


def _modules():
    """sys.modules entries standing in for tensorflow / deepctr while the reference's nets code is imported and run."""
    layer_ns = {k: v for k, v in globals().items() if isinstance(v, type) and issubclass(v, Layer) and v is not Layer}
    layer_ns["Input"] = Input
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)       # importlib.util.find_spec(name) must not raise
        m.__dict__.update(attrs)
        return m

    layers = mod("tensorflow.keras.layers", **layer_ns)
    backend = mod("tensorflow.keras.backend",
                  get_session=lambda: types.SimpleNamespace(run=lambda *a, **k: None))
    keras = mod("tensorflow.keras", layers=layers, regularizers=mod("tensorflow.keras.regularizers"),
                models=mod("tensorflow.keras.models", Model=Model), backend=backend)
    tf = mod("tensorflow", keras=keras,
             fill=lambda dims, value: np.full(tuple(int(d) for d in dims), value),
             shape=lambda x: np.asarray(x).shape,
             squeeze=lambda x, axis=None: np.squeeze(x, axis=axis),
             math=types.SimpleNamespace(reduce_mean=lambda x, axis=None, keepdims=False: np.mean(x, axis=axis, keepdims=keepdims)),
             global_variables_initializer=lambda: None)
    seq = mod("deepctr.layers.sequence", AttentionSequencePoolingLayer=AttentionSequencePoolingLayer, DynamicGRU=DynamicGRU)
    dl = mod("deepctr.layers", sequence=seq)
    return {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.regularizers": keras.regularizers, "tensorflow.keras.models": keras.models,
            "tensorflow.keras.backend": backend, "deepctr": mod("deepctr", layers=dl), "deepctr.layers": dl,
            "deepctr.layers.sequence": seq}


@contextlib.contextmanager
def _swapped_modules():
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    ours = _modules()
    stale = [k for k in sys.modules if k == "rl4rs.nets" or k.startswith("rl4rs.nets.")]
    saved = {k: sys.modules.get(k) for k in list(ours) + stale + ["rl4rs"]}
    try:
        for k in stale:
            del sys.modules[k]
        sys.modules.update(ours)
        if "rl4rs" not in sys.modules:                   # the package __init__ registers gym ids: bypass it
            pkg = types.ModuleType("rl4rs")
            pkg.__path__ = [REFERENCE_ROOT + "/rl4rs"]
            sys.modules["rl4rs"] = pkg
        yield
    finally:
        for k in [k for k in sys.modules if k == "rl4rs.nets" or k.startswith("rl4rs.nets.")]:
            del sys.modules[k]                           # they hold references to the stand-ins
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_reference_graph(algo, config, checkpoint, seq, dense, cat):
    """rl4rs.nets.<algo>.get_model(config) (slate.py:239-242) on the fed rows.
    -> dict(obs = output of the layer named simulator_obs, probs = simulator_reward (slate.py:228-237),
            variables = [(top-level scope, variable name, shape)] in creation order, unused = checkpoint entries the
            graph never asked for, layers = every layer name in construction order)."""
    global _CTX
    feed = {"sequence_feature_input": np.asarray(seq), "dense_feature_input": np.asarray(dense),
            "category_feature_input": np.asarray(cat), "slate_label": np.zeros((len(cat), 9), np.int64)}
    _CTX = _Ctx(feed, checkpoint)
    try:
        with _swapped_modules():
            model = importlib.import_module("rl4rs.nets." + algo).get_model(config)
        pick = lambda key: [l for l in model.layers if key in l.name][0]                     # slate.py:230-236
        return {"obs": pick("simulator_obs").output, "probs": pick("simulator_reward").output,
                "variables": list(_CTX.created), "unused": sorted(set(_CTX.ckpt) - _CTX.used),
                "layers": [l.name for l in model.layers]}
    finally:
        _CTX = None


# ---- the reference's whole stack: RecEnvBase -> SlateRecEnv.__init__ -> RecSimBase.__init__ -> get_model -> Saver.restore ----
class _Scope(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Graph(object):
    """tf.Graph(): a fresh graph is a fresh Keras name space (base.py:119) -- layer uids restart here."""

    def __init__(self):
        global _CTX
        _CTX = _Ctx(None, None)

    def as_default(self):
        return _Scope()


class _Session(object):
    def __init__(self, graph=None, config=None):
        self.graph = graph

    def as_default(self):
        return _Scope()


class _Stack(object):
    """What the TF runtime would hold for one simulator: the config its graph was built from and the restored values."""

    def __init__(self, config):
        self.config, self.ckpt, self.restored_from, self.calls = config, None, None, 0


class _Saver(object):
    def __init__(self, stack):
        self.stack = stack

    def restore(self, sess, model_file):            # base.py:148-151: a tf.train.Saver prefix, read WITHOUT TensorFlow
        from rl4rs_b200.utils.tf_checkpoint import TensorBundleReader
        rd = TensorBundleReader(model_file)
        self.stack.ckpt = {name: np.asarray(rd.get_tensor(name), DT) for name in rd.variables()}    # converted once
        self.stack.restored_from = model_file


class _Function(object):
    """tf.keras.backend.function(model.input, layer.output) (slate.py:232-237): evaluating it re-runs the reference's
    get_model eagerly on the fed rows with the restored variables and hands back that layer's output as float32."""

    def __init__(self, stack, output):
        name = _CTX.requested.get(id(output))
        assert name is not None and [l for l in _CTX.layers if l.name == name][0].output is output
        self.stack, self.layer, self.inputs = stack, name, list(_CTX.inputs)

    def __call__(self, feat):
        global _CTX
        st = self.stack
        assert st.ckpt is not None, "Saver.restore has not run"
        prev, _CTX = _CTX, _Ctx(dict(zip(self.inputs, feat)), st.ckpt)
        try:
            algo = st.config.get("algo", "dien")
            model = importlib.import_module("rl4rs.nets." + algo).get_model(st.config)
            assert not (set(_CTX.ckpt) - _CTX.used), sorted(set(_CTX.ckpt) - _CTX.used)
            st.calls += 1
            return np.asarray(model.get_layer(self.layer).output, np.float32)
        finally:
            _CTX = prev


@contextlib.contextmanager
def full_reference_stack(config):
    """Inside this context `SlateRecEnv(config, SlateState)` / `SeqSlateRecEnv(...)` of the REFERENCE construct themselves
    with their own __init__ (base.py:114-131, slate.py:223-237): tf.Graph / Session / train.Saver / keras.backend.function
    are the stand-ins above, `config['model_file']` must be a Saver prefix (rl4rs_b200.utils.tf_checkpoint writes one).
    Yields (ref_base, ref_slate, ref_seqslate, stack)."""
    global _CTX
    from oracle import ref_harness
    ref_base, ref_slate, ref_seqslate, _ = ref_harness.install_stubs()
    stack = _Stack(config)
    tfm = ref_base.tf
    assert tfm is ref_slate.tf
    patched = {"Graph": _Graph, "Session": _Session, "ConfigProto": lambda **kw: None,
               "train": types.SimpleNamespace(Saver=lambda: _Saver(stack))}
    saved = {k: getattr(tfm, k, None) for k in patched}
    saved_backend = getattr(tfm.keras, "backend", None)
    try:
        for k, v in patched.items():
            setattr(tfm, k, v)
        tfm.keras.backend = types.SimpleNamespace(function=lambda inputs, output: _Function(stack, output))
        with _swapped_modules():
            yield ref_base, ref_slate, ref_seqslate, stack
    finally:
        for k, v in saved.items():
            if v is None:
                delattr(tfm, k)
            else:
                setattr(tfm, k, v)
        if saved_backend is None:
            del tfm.keras.backend
        else:
            tfm.keras.backend = saved_backend
        _CTX = None


# ---- the RLlib policy models (rl4rs/nets/rllib/rllib_rawstate_model.py:25-86, rllib_mask_model.py:7-64) -----------------
class _TFModelV2(object):
    def __init__(self, obs_space, action_space, num_outputs, model_config, name, *a, **kw):
        self.obs_space, self.action_space, self.num_outputs = obs_space, action_space, num_outputs
        self.model_config, self.name = model_config, name


class _ParametricActionsModel(_TFModelV2):
    """ray.rllib.examples.models.parametric_actions_model.ParametricActionsModel: owns `action_embed_model`, an RLlib
    FullyConnectedNetwork over the true observation.  Here that sub-model is whatever the test plugs in."""

    def __init__(self, obs_space, action_space, num_outputs, model_config, name, true_obs_shape=None, action_embed_size=None, **kw):
        super().__init__(obs_space, action_space, num_outputs, model_config, name)
        self.action_embed_model = types.SimpleNamespace(model_config=model_config)


@contextlib.contextmanager
def _rllib_stand_ins():
    """sys.modules entries for the ray / tensorflow names the two RLlib model files import; yields the tf stand-in."""
    mods = _modules()
    tf = mods["tensorflow"]
    tf.keras.Model = Model
    tf.reshape = lambda x, shape: np.reshape(x, shape)
    tf.maximum = np.maximum
    def _log(x):
        with np.errstate(divide="ignore"):            # log(0) = -inf is the point of the mask (rllib_mask_model.py:61)
            return np.log(np.asarray(x, np.float32))
    tf.math.log = _log
    tf.float32 = types.SimpleNamespace(min=np.finfo(np.float32).min)
    Model.summary = lambda self: None

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__dict__.update(attrs)
        return m

    ray = {"ray": module("ray"), "ray.rllib": module("ray.rllib"), "ray.rllib.models": module("ray.rllib.models"),
           "ray.rllib.models.utils": module("ray.rllib.models.utils", get_activation_fn=lambda name, framework="tf": None),
           "ray.rllib.models.tf": module("ray.rllib.models.tf"),
           "ray.rllib.models.tf.misc": module("ray.rllib.models.tf.misc", normc_initializer=lambda std=1.0: None),
           "ray.rllib.models.tf.tf_modelv2": module("ray.rllib.models.tf.tf_modelv2", TFModelV2=_TFModelV2),
           "ray.rllib.utils": module("ray.rllib.utils"),
           "ray.rllib.utils.framework": module("ray.rllib.utils.framework", try_import_tf=lambda: (tf, tf, 1),
                                               try_import_torch=lambda: (None, None)),
           "ray.rllib.examples": module("ray.rllib.examples"), "ray.rllib.examples.models": module("ray.rllib.examples.models"),
           "ray.rllib.examples.models.parametric_actions_model":
               module("ray.rllib.examples.models.parametric_actions_model", ParametricActionsModel=_ParametricActionsModel)}
    saved = {k: sys.modules.get(k) for k in ray}
    try:
        with _swapped_modules():
            sys.modules.update(ray)
            sys.modules["tensorflow"] = tf                # the module try_import_tf hands out IS the stand-in
            yield tf
    finally:
        for k in [k for k in sys.modules if k.startswith("rl4rs.nets")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_reference_rawstate_model(config, checkpoint, category, dense, sequence, num_outputs):
    """Builds the reference's own TFModelWithRawState on the fed raw-state rows (its Keras graph is assembled -- and,
    here, evaluated -- in __init__) under stand-ins for the ray / gym names it imports.
    -> dict(logits = 'fc_out' (before the action mask), value = 'value_out', variables, layers)."""
    global _CTX
    from oracle import ref_harness
    ref_harness.install_stubs()
    gym = sys.modules["gym"]

    class Space(object):
        def __init__(self, shape):
            self.shape = tuple(shape)

    class DictSpace(gym.spaces.Dict):                    # isinstance(obs_space, gym.spaces.Dict) + item access + .original_space
        def __getitem__(self, key):
            return self.spaces[key]

        @property
        def original_space(self):
            return self

    obs_space = DictSpace({"category_feature": Space(np.shape(category)[1:]), "dense_feature": Space(np.shape(dense)[1:]),
                           "sequence_feature": Space(np.shape(sequence)[1:])})
    feed = {"obs_category_input": np.asarray(category), "obs_dense_input": np.asarray(dense),
            "obs_sequence_input": np.asarray(sequence)}
    _CTX = _Ctx(feed, checkpoint)
    try:
        with _rllib_stand_ins():
            mod = importlib.import_module("rl4rs.nets.rllib.rllib_rawstate_model")
            model = mod.TFModelWithRawState(obs_space, None, num_outputs, {}, "rawstate", config)
            by_name = {l.name: l for l in _CTX.layers}
            return {"logits": by_name["fc_out"].output, "value": np.reshape(by_name["value_out"].output, [-1]),
                    "variables": list(_CTX.created), "unused": sorted(set(_CTX.ckpt) - _CTX.used),
                    "layers": [l.name for l in _CTX.layers], "model": type(model).__name__}
    finally:
        _CTX = None


def run_reference_mask_forward(action_embed_fn, obs, action_mask, action_size):
    """MyMaskActionsModel.forward (rllib_mask_model.py:41-62) of the reference on a batch: `action_embed_fn(obs)` stands
    for RLlib's FullyConnectedNetwork (the `action_embed_model` the parent class owns); what is the reference's own is
    the masking  logits + max(log(action_mask), float32.min).  -> masked logits."""
    global _CTX
    from oracle import ref_harness
    ref_harness.install_stubs()
    _CTX = _Ctx(None, None)
    try:
        with _rllib_stand_ins():
            mod = importlib.import_module("rl4rs.nets.rllib.rllib_mask_model")
            cls = mod.getMaskActionsModel((np.shape(obs)[1],), action_size)
            model = cls(None, None, action_size, {}, "mask_model")
            assert model.model_config["fcnet_hiddens"] == [64] and model.model_config["vf_share_layers"] is True
            model.action_embed_model = lambda d: (action_embed_fn(d["obs"]), None)
            out, state = model.forward({"obs": {"obs": obs, "action_mask": np.asarray(action_mask, np.float32)}}, [], None)
            return np.asarray(out)
    finally:
        _CTX = None
