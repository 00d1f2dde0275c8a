"""numpy stand-ins for the third-party names the reference's serving path imports.  TEST INFRASTRUCTURE ONLY.

TensorFlow, Keras, tensorflow_recommenders and faiss are not installed in this image, so the reference's
own files cannot be imported as they are.  `install()` registers minimal numpy modules under those four
names -- only the names `src/models.py`, `src/trainer.py:219-246` and `app/recommendation_service.py` touch
on the serving path -- so that those files run UNMODIFIED from /root/reference (make_reference_golden.py).
What that pins: the reference's own composition (layer order, concat order, the cross formula with the bias
outside the product, which tensors feed the heads, the recommend / score / cold-start control flow, the
catalog build).  What it does not pin: the primitives' arithmetic, which is this file's numpy statement of
the documented op (Dense = x @ kernel + bias, Embedding = row gather, StringLookup OOV -> 0, IndexFlatIP =
inner product, top-k by score descending and row ascending, -1 / -FLT_MAX padding), not TF's or faiss's code.

The arithmetic type is `set_compute_dtype(np.float64 | np.float32)`; fp64 is the yardstick the fixtures keep.
Anything outside that list (tf.data, optimizers, losses ...) is a placeholder that imports but cannot compute.
"""

import sys
import types

import numpy as np

_DTYPE = [np.float64]
NEG_PAD = np.float32(-3.4028234663852886e38)   # faiss's neutral heap value for inner product


def set_compute_dtype(dt):
    _DTYPE[0] = dt


class Tensor(np.ndarray):
    """An ndarray with `.numpy()`; operators keep the subclass, which is all the reference needs of a tf.Tensor."""

    def numpy(self):
        return np.array(self, copy=True).view(np.ndarray)


def _t(a):
    return np.asarray(a).view(Tensor)


class _Placeholder:
    """Importable, callable, attribute-able; any arithmetic on it fails.  For names off the serving path."""

    def __init__(self, name="placeholder"):
        self._name = name

    def __call__(self, *a, **k):
        return _Placeholder(self._name + "()")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Placeholder(f"{self._name}.{item}")


def _shape_of(x):
    if isinstance(x, dict):
        return {k: _shape_of(v) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return [_shape_of(v) for v in x]
    return tuple(np.shape(x))


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.built = False

    def add_weight(self, name=None, shape=None, initializer="zeros", **kwargs):
        return _t(np.zeros(shape, dtype=np.float64))   # values are assigned by the caller afterwards

    def build(self, input_shape):
        pass

    def call(self, inputs, **kwargs):
        raise NotImplementedError

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(_shape_of(inputs))
            self.built = True
        return self.call(inputs, *args, **kwargs)

    def get_config(self):
        return {"name": self.name}


class Model(Layer):
    _saved = {}

    def save(self, path, **kwargs):
        """Keeps the object under its path (stands for the .keras archive) and leaves a marker file there."""
        Model._saved[str(path)] = self
        with open(path, "wb") as f:
            f.write(b"stand-in for a .keras archive; see tests/golden/tf_faiss_standin.py\n")


def load_model(path, custom_objects=None, **kwargs):
    return Model._saved[str(path)]


class Dense(Layer):
    """keras.layers.Dense: activation(inputs @ kernel + bias), kernel stored (in, out)."""

    def __init__(self, units, activation=None, kernel_regularizer=None, name=None, **kwargs):
        super().__init__(name=name)
        self.units, self.activation = units, activation
        self.kernel = self.bias = None

    def build(self, input_shape):
        self.kernel = np.zeros((input_shape[-1], self.units))
        self.bias = np.zeros((self.units,))

    def call(self, inputs, training=None):
        dt = _DTYPE[0]
        y = np.asarray(inputs, dtype=dt) @ self.kernel.astype(dt) + self.bias.astype(dt)
        if self.activation == "relu":
            y = np.maximum(y, 0)
        elif self.activation == "sigmoid":
            y = 1.0 / (1.0 + np.exp(-y))
        elif self.activation is not None:
            raise ValueError(self.activation)
        return _t(y.astype(dt))


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, **kwargs):
        super().__init__(**kwargs)
        self.input_dim, self.output_dim = input_dim, output_dim
        self.embeddings = np.zeros((input_dim, output_dim))

    def call(self, inputs, training=None):
        return _t(self.embeddings[np.asarray(inputs)].astype(_DTYPE[0]))


class StringLookup(Layer):
    """keras.layers.StringLookup(vocabulary=v, mask_token=None): one OOV bucket at 0, v[i] -> i + 1, int64."""

    def __init__(self, vocabulary=None, mask_token=None, num_oov_indices=1, **kwargs):
        super().__init__(**kwargs)
        assert mask_token is None and num_oov_indices == 1
        self._index = {s: i + 1 for i, s in enumerate(vocabulary)}

    def call(self, inputs, training=None):
        flat = [s.decode() if isinstance(s, bytes) else str(s) for s in np.asarray(inputs).ravel()]
        return _t(np.array([self._index.get(s, 0) for s in flat], dtype=np.int64).reshape(np.shape(inputs)))


class Sequential(Model):
    def __init__(self, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.layers = list(layers or [])

    def call(self, inputs, training=None):
        x = inputs
        for layer in self.layers:
            x = layer(x)
        return x


class IndexFlatIP:
    """faiss.IndexFlatIP: exact inner-product search.  Scores in the compute dtype, reported as float32; ordered by
    (score descending, row ascending); lists shorter than k are padded with label -1 and -FLT_MAX."""

    def __init__(self, d):
        self.d = int(d)
        self._xb = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self):
        return self._xb.shape[0]

    def add(self, x):
        assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2 and x.shape[1] == self.d
        self._xb = np.vstack([self._xb, x])

    def search(self, x, k):
        assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2 and x.shape[1] == self.d
        dt = _DTYPE[0]
        s = x.astype(dt) @ self._xb.astype(dt).T
        nq, n = s.shape
        D = np.full((nq, k), NEG_PAD, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        kk = min(k, n)
        for qi in range(nq):
            order = np.lexsort((np.arange(n), -s[qi]))[:kk]
            I[qi, :kk] = order
            D[qi, :kk] = s[qi, order].astype(np.float32)
        return D, I


_written = {}


def write_index(index, path):
    _written[str(path)] = index
    with open(path, "wb") as f:
        f.write(b"stand-in for a faiss index file; see tests/golden/tf_faiss_standin.py\n")


def read_index(path):
    return _written[str(path)]


def normalize_L2(x):
    """faiss.normalize_L2: in place, each row scaled by 1 / sqrt(sum of squares); all-zero rows are left alone."""
    assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2
    dt = _DTYPE[0]
    nr = (x.astype(dt) ** 2).sum(axis=1)
    scale = np.where(nr > 0, 1.0 / np.sqrt(np.where(nr > 0, nr, 1.0)), 1.0)
    x[...] = (x.astype(dt) * scale[:, None]).astype(np.float32)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def fallback(item, _n=name):   # PEP 562: every other public name imports as a placeholder
        if item.startswith("__"):
            raise AttributeError(item)
        return _Placeholder(f"{_n}.{item}")

    m.__getattr__ = fallback
    return m


def install():
    """Register the stand-in modules (idempotent).  Returns the dict of modules it put into sys.modules."""
    layers = _module("tensorflow.keras.layers", Layer=Layer, Dense=Dense, Embedding=Embedding,
                     StringLookup=StringLookup)
    utils = _module("tensorflow.keras.utils", register_keras_serializable=lambda *a, **k: (lambda cls: cls))
    models = _module("tensorflow.keras.models", load_model=load_model, Model=Model, Sequential=Sequential)
    callbacks = _module("tensorflow.keras.callbacks", Callback=type("Callback", (), {"__init__": lambda self: None}))
    keras = _module("tensorflow.keras", Model=Model, Sequential=Sequential, layers=layers, utils=utils,
                    models=models, callbacks=callbacks)
    linalg = _module("tensorflow.linalg", matvec=lambda a, b: _t(
        np.asarray(a, dtype=_DTYPE[0]) @ np.asarray(b, dtype=_DTYPE[0])))
    tf = _module(
        "tensorflow", keras=keras, linalg=linalg,
        constant=lambda v, dtype=None: _t(np.asarray(v)),
        expand_dims=lambda x, axis: _t(np.expand_dims(np.asarray(x), axis)),
        squeeze=lambda x, axis=None: _t(np.squeeze(np.asarray(x), axis=axis)),
        concat=lambda xs, axis: _t(np.concatenate([np.asarray(v) for v in xs], axis=axis)),
        tensordot=lambda a, b, axes: _t(np.tensordot(np.asarray(a, dtype=_DTYPE[0]),
                                                     np.asarray(b, dtype=_DTYPE[0]), axes=axes)),
    )
    tfrs = _module("tensorflow_recommenders",
                   models=_module("tensorflow_recommenders.models", Model=Model),
                   tasks=_module("tensorflow_recommenders.tasks"))
    faiss = _module("faiss", IndexFlatIP=IndexFlatIP, Index=IndexFlatIP, normalize_L2=normalize_L2,
                    write_index=write_index, read_index=read_index)
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.utils": utils, "tensorflow.keras.models": models,
            "tensorflow.keras.callbacks": callbacks, "tensorflow.linalg": linalg,
            "tensorflow_recommenders": tfrs, "tensorflow_recommenders.models": tfrs.models,
            "tensorflow_recommenders.tasks": tfrs.tasks, "faiss": faiss}
    sys.modules.update(mods)
    return mods


def uninstall():
    for name in [n for n in sys.modules if n.split(".")[0] in ("tensorflow", "tensorflow_recommenders", "faiss")]:
        if getattr(sys.modules[name], "__file__", None) is None:
            del sys.modules[name]


def import_reference(root="/root/reference"):
    """Import src.config, src.models, src.evaluation, src.trainer and app.recommendation_service from the files
    where they lie, bypassing the two package __init__ files (src/__init__.py pulls in the whole training stack
    eagerly; nothing in it is on the serving path)."""
    import importlib
    from pathlib import Path
    root = Path(root)
    for pkg in ("src", "app"):
        m = types.ModuleType(pkg)
        m.__path__ = [str(root / pkg)]
        sys.modules[pkg] = m
    out = {}
    for name in ("src.config", "src.models", "src.evaluation", "app.recommendation_service"):
        out[name] = importlib.import_module(name)
    try:
        out["src.trainer"] = importlib.import_module("src.trainer")
    except Exception as e:   # its module-level imports reach past the stand-ins (tf.data, sklearn ...)
        out["src.trainer"] = None
        out["src.trainer.error"] = repr(e)
    return out


def drop_reference(root="/root/reference"):
    for name in [n for n in sys.modules if n == "src" or n == "app" or n.startswith("src.") or n.startswith("app.")]:
        del sys.modules[name]
    sys.path[:] = [p for p in sys.path if p != str(root)]          # app/main.py:18 puts the reference root on sys.path
