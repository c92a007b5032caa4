"""An EAGER stand-in for the ~60 `tensorflow.compat.v1` / `tensorflow_probability` calls the reference makes.

TEST INFRASTRUCTURE.  Purpose: let the reference's OWN source files (/root/reference/vcsmc.py, vncsmc.py) be
imported UNMODIFIED and executed line by line in the build container, where TensorFlow 1.15 / TFP 0.7 cannot be
installed (Python 3.12, no network), so that golden vectors for the whole SMC step come from the reference's code
and not from a human restatement of it.  tests/golden/make_golden.py drives it; nothing else imports it, and
nothing in the test-suite needs it at run time (the tests read the committed .npz files).

How it works
  * every `tf.*` function below is executed immediately on torch-CPU float64 tensors (string tensors are numpy
    object arrays), so "building the graph" (`VCSMC.sample_phylogenies()`) IS running the sweep;
  * `tf.placeholder` returns the array handed to `feed()` beforehand (the reference feeds `self.core`);
  * `tf.Variable` makes a torch leaf that requires grad, so `torch.autograd.grad(model.cost, variables())` is what
    `optimizer.minimize(self.cost)` differentiates (vcsmc.py:488-491); `variable_overrides[name]` replaces an
    initial value (same effect as a `tf.assign` before the first `sess.run`);
  * randomness is INJECTED: `tf.random.uniform`, `tf.random.categorical` and `tfp.distributions.Exponential.sample`
    pop explicit uniform arrays from `push_uniforms(...)` in call order.  The three samplers follow the published
    algorithms of tensorflow==1.15.0 / tensorflow_probability==0.7.0 (requirements.txt:2-3):
      - categorical: per row, running float64 sum of exp(logit - rowmax) in column order; draw = first column whose
        running sum exceeds u * total (upper bound), clamped to the last column;
      - Exponential.sample: -log(U)/rate with U in [tiny, 1), reparameterised (differentiable in the rate);
      - uniform: the injected float32 array as is;
  * `tf.nn.top_k`: descending, ties towards the LOWER index (stable);
  * `tf.linalg.expm`: scipy.linalg.expm (Pade 13 + scaling and squaring, the algorithm TF implements) with the
    adjoint from the block-triangular identity expm([[A^T, G], [0, A^T]])[:n, n:] = L(A^T, G);
  * every call of the samplers and of top_k is logged in `trace` so the integer decisions (ancestors, pairs,
    remaining order, nested choices) can be stored next to the floating-point outputs.

dtype rules kept from TF: python floats / numpy float64 -> float64, python ints -> int32, `int / int` -> float64
(true division), float32 only where the reference asks for it (`tf.random.uniform`, `tf.cast(.., tf.float32)`,
`tf.one_hot`).
"""
import builtins as _b
import sys
import types

import numpy as np
import scipy.linalg as spl
import torch

float64 = torch.float64
float32 = torch.float32
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001  (tf.bool)

_py_bool = _b.bool
_range = _b.range

trace = {"categorical": [], "top_k": [], "uniform": [], "exponential": []}
variable_overrides = {}
_variables = []
_uniform_queue = []
_feed = [None]


# ----------------------------------------------------------------------------------------------------------
# driver-side controls
# ----------------------------------------------------------------------------------------------------------
def reset():
    for v in trace.values():
        v.clear()
    variable_overrides.clear()
    _variables.clear()
    _uniform_queue.clear()
    _feed[0] = None


def feed(array):
    _feed[0] = torch.as_tensor(np.asarray(array, dtype=np.float64))


def push_uniforms(kind, array):
    """kind in {'uniform','categorical','exponential'}: checked against the consumer, so that a change in the
    reference's order of random calls cannot go unnoticed."""
    _uniform_queue.append((kind, np.asarray(array)))


def _pop(kind, shape):
    assert _uniform_queue, "the reference asked for more random numbers than were injected (%s %s)" % (kind, shape)
    k, a = _uniform_queue.pop(0)
    assert k == kind, "random call order: reference asks for %s, injected %s" % (kind, k)
    assert tuple(a.shape) == tuple(shape), "random shape: reference asks for %s %s, injected %s" % (kind, shape, a.shape)
    return a


def uniforms_left():
    return len(_uniform_queue)


def variables():
    return list(_variables)


# ----------------------------------------------------------------------------------------------------------
# conversions
# ----------------------------------------------------------------------------------------------------------
def _is_str(x):
    return isinstance(x, np.ndarray) and x.dtype == object


def _t(x, dtype=None):
    """Anything -> torch tensor (or numpy object array for strings), TF auto-packing rules."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if _is_str(x):
        return x
    if isinstance(x, (list, tuple)):
        if len(x) and all(isinstance(e, str) or (isinstance(e, (list, tuple)) and len(e) and isinstance(e[0], str)) for e in x):
            return np.array(x, dtype=object)
        if any(isinstance(e, (torch.Tensor, list, tuple)) for e in x):
            parts = [_t(e, dtype) for e in x]
            if dtype is None:
                dt = parts[0].dtype
                for p in parts[1:]:
                    dt = torch.promote_types(dt, p.dtype)
                parts = [p.to(dt) for p in parts]
            return torch.stack(parts)
    if isinstance(x, str):
        return np.array(x, dtype=object)
    a = np.asarray(x)
    if a.dtype == np.int64 and not isinstance(x, np.ndarray):
        a = a.astype(np.int32)                     # python ints are int32 in TF
    out = torch.as_tensor(a)
    return out if dtype is None else out.to(dtype)


def _i(x):
    return int(x)


def _ints(seq):
    if isinstance(seq, torch.Tensor):
        return [int(v) for v in seq.reshape(-1)]
    return [int(v) for v in seq]


def _idx(x):
    return _t(x).to(torch.int64)


# ----------------------------------------------------------------------------------------------------------
# tf.* functions (only what vcsmc.py / vncsmc.py call)
# ----------------------------------------------------------------------------------------------------------
def Variable(initial_value, dtype=None, name=None):
    init = variable_overrides.get(name, initial_value)
    v = torch.as_tensor(np.array(init, dtype=np.float64)).clone().requires_grad_(True)
    v._shim_name = name
    _variables.append(v)
    return v


def constant(value, dtype=None, shape=None, name=None):
    if isinstance(value, str) or (isinstance(value, (list, tuple)) and _is_str(_t(value))):
        a = _t(value)
        if shape is not None:
            a = np.full(tuple(shape), value, dtype=object) if isinstance(value, str) else a.reshape(shape)
        return a
    t = _t(value, dtype)
    if shape is not None:
        t = t.expand(*_ints(shape)).clone() if t.dim() == 0 else t.reshape(*_ints(shape))
    return t


def placeholder(dtype=None, shape=None, name=None):
    assert _feed[0] is not None, "feed() the placeholder before building"
    return _feed[0]


def convert_to_tensor(value=None, dtype=None):
    return _t(value, dtype)


def cast(x, dtype):
    return _t(x).to(dtype)


def exp(x):
    return torch.exp(_t(x))


def log(x):
    x = _t(x)
    if x.dtype == torch.float32 and not x.requires_grad:      # float32 logs via numpy (what the oracle and the
        with np.errstate(divide="ignore", invalid="ignore"):  # CUDA tie rule were written against)
            return torch.from_numpy(np.log(x.numpy()).astype(np.float32))
    return torch.log(x)


def negative(x):
    return -_t(x)


def multiply(a, b):
    return _t(a) * _t(b)


def maximum(a, b):
    a = _t(a)
    return torch.maximum(a, _t(b).to(a.dtype))


def equal(a, b):
    return _t(a) == _t(b)


def greater_equal(a, b):
    return _t(a) >= _t(b)


def where(c, a, b):
    return torch.where(c, a, b)


def ones_like(x):
    return torch.ones_like(_t(x))


def mod(a, b):
    return torch.remainder(_t(a), _i(b) if not isinstance(b, torch.Tensor) or b.dim() == 0 else b)


def floordiv(a, b):
    return torch.div(_t(a), _i(b), rounding_mode="floor")


def shape(x):
    x = _t(x)
    return torch.tensor(list(x.shape), dtype=torch.int32)


def range(*args):  # noqa: A001
    return torch.arange(*[_i(a) for a in args], dtype=torch.int32)


def reshape(x, shp):
    x = _t(x)
    shp = _ints(shp)
    if -1 in shp and 0 in shp:
        # TF's Reshape kernel leaves zero-sized dimensions out of both products when it infers the -1 entry
        have = int(np.prod([d for d in x.shape if d != 0], dtype=np.int64))
        want = int(np.prod([d for d in shp if d not in (0, -1)], dtype=np.int64))
        shp = [have // want if d == -1 else d for d in shp]
    return x.reshape(shp)


def transpose(x, perm=None):
    x = _t(x)
    if _is_str(x):
        return x.T if perm is None else np.transpose(x, perm)
    if perm is None:
        return x.permute(*reversed(_range(x.dim())))
    return x.permute(*_ints(perm))


def expand_dims(x, axis):
    x = _t(x)
    if _is_str(x):
        return np.expand_dims(x, axis)
    return x.unsqueeze(axis)


def squeeze(x, axis=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def tile(x, multiples):
    x = _t(x)
    return x.repeat(*_ints(multiples))


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def concat(values, axis):
    parts = [v for v in values]
    first = next((p for p in parts if isinstance(p, torch.Tensor) or _is_str(p)), None)
    if _is_str(first):
        return np.concatenate([_t(p) for p in parts], axis=axis)
    dt = first.dtype if first is not None else None
    return torch.cat([_t(p, dt) for p in parts], dim=axis)   # list elements are cast to the tensors' dtype


def gather(params, indices, axis=0):
    params = _t(params)
    assert axis == 0
    if _is_str(params):
        return params[_idx(indices).numpy()]
    return params[_idx(indices)]


def gather_nd(params, indices):
    params, ix = _t(params), _idx(indices)
    return params[tuple(ix[..., d] for d in _range(ix.shape[-1]))]


def one_hot(indices, depth):
    return torch.nn.functional.one_hot(_idx(indices), _i(depth)).to(torch.float32)


def _axis(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (tuple, list)) else axis


def reduce_sum(x, axis=None):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axis(axis))


def reduce_mean(x, axis=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axis(axis))


def reduce_prod(x, axis=None):
    x = _t(x)
    return x.prod().to(x.dtype) if axis is None else x.prod(dim=axis)


def reduce_logsumexp(x, axis=None):
    x = _t(x)
    return torch.logsumexp(x.reshape(-1), dim=0) if axis is None else torch.logsumexp(x, dim=axis)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def tensordot(a, b, axes):
    a, b = _t(a), _t(b)
    assert axes == 0
    return a.reshape(*a.shape, *([1] * b.dim())) * b


def einsum(eq, *ops):
    return torch.einsum(eq, *[_t(o) for o in ops])


def cond(pred, true_fn, false_fn):
    return true_fn() if _py_bool(pred) else false_fn()


def while_loop(cond, body, loop_vars, shape_invariants=None):  # noqa: A002
    vs = list(loop_vars)
    while _py_bool(cond(*vs)):
        vs = list(body(*vs))
    return vs


class TensorShape(list):
    pass


def get_variable_scope():
    return types.SimpleNamespace(name="")


def get_collection(*a, **k):
    return variables()


def global_variables_initializer():
    return None


def device(_):
    import contextlib
    return contextlib.nullcontext()


# --- tf.math -----------------------------------------------------------------------------------------------
math = types.ModuleType("tensorflow.compat.v1.math")
math.log = log


def _count_nonzero(x):
    return torch.count_nonzero(_t(x))


math.count_nonzero = _count_nonzero


# --- tf.linalg ---------------------------------------------------------------------------------------------
class _Expm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        ctx.save_for_backward(A)
        return torch.from_numpy(np.asarray(spl.expm(A.detach().numpy())))

    @staticmethod
    def backward(ctx, G):
        (A,) = ctx.saved_tensors
        a = A.detach().numpy()
        n = a.shape[-1]
        at = np.swapaxes(a, -1, -2)
        blk = np.zeros(a.shape[:-2] + (2 * n, 2 * n))
        blk[..., :n, :n] = at
        blk[..., n:, n:] = at
        blk[..., :n, n:] = G.numpy()
        return torch.from_numpy(np.asarray(spl.expm(blk))[..., :n, n:].copy())


def _set_diag(x, diagonal):
    x = _t(x)
    d = _t(diagonal).to(x.dtype)
    n = x.shape[-1]
    eye = torch.eye(n, dtype=x.dtype)
    return x * (1 - eye) + torch.diag_embed(d.expand(*x.shape[:-1]))


linalg = types.ModuleType("tensorflow.compat.v1.linalg")
linalg.expm = lambda A: _Expm.apply(_t(A))
linalg.set_diag = _set_diag


# --- tf.random ---------------------------------------------------------------------------------------------
def _random_uniform(shp, minval=0, maxval=1, dtype=float32):
    shp = _ints(shp)
    assert minval == 0 and maxval == 1
    u = _pop("uniform", shp).astype(np.float32)
    trace["uniform"].append(u.copy())
    return torch.from_numpy(u.copy())


def _random_categorical(logits, num_samples):
    lg = _t(logits).detach().numpy().astype(np.float64)        # [B, C]
    B, C = lg.shape
    n = _i(num_samples)
    u = _pop("categorical", (B, n)).astype(np.float64)
    w = np.exp(lg - lg.max(axis=1, keepdims=True))
    cdf = np.cumsum(w, axis=1)                                 # running double sum in column order
    out = np.empty((B, n), dtype=np.int64)
    for b in _range(B):
        out[b] = np.minimum(np.searchsorted(cdf[b], u[b] * cdf[b, -1], side="right"), C - 1)
    trace["categorical"].append(out.copy())
    return torch.from_numpy(out)


random = types.ModuleType("tensorflow.compat.v1.random")
random.uniform = _random_uniform
random.categorical = _random_categorical


# --- tf.nn -------------------------------------------------------------------------------------------------
def _top_k(x, k):
    x = _t(x)
    a = x.detach().numpy()
    k = _i(k)
    order = np.argsort(-a, axis=-1, kind="stable")[..., :k]   # descending, lower index first on ties
    vals = np.take_along_axis(a, order, axis=-1)
    trace["top_k"].append(order.astype(np.int32).copy())
    return torch.from_numpy(vals.copy()), torch.from_numpy(order.astype(np.int32))


nn = types.ModuleType("tensorflow.compat.v1.nn")
nn.top_k = _top_k


# --- tfp ---------------------------------------------------------------------------------------------------
class _Exponential:
    def __init__(self, rate):
        self.rate = _t(rate)

    def sample(self, n):
        n = _i(n)
        u = _pop("exponential", (n,)).astype(np.float64)
        trace["exponential"].append(u.copy())
        return -torch.log(torch.from_numpy(u.copy())) / self.rate


# ----------------------------------------------------------------------------------------------------------
# installation
# ----------------------------------------------------------------------------------------------------------
def install():
    """Registers this module as `tensorflow.compat.v1`, plus `tensorflow_probability` and the plotting stubs."""
    torch.set_default_dtype(torch.float64)      # int / int -> float64 like TF's true division
    if not hasattr(torch.Tensor, "get_shape"):
        torch.Tensor.get_shape = lambda self: self.shape
    me = sys.modules[__name__]
    tf_root = types.ModuleType("tensorflow")
    compat = types.ModuleType("tensorflow.compat")
    tf_root.compat = compat
    compat.v1 = me
    sys.modules["tensorflow"] = tf_root
    sys.modules["tensorflow.compat"] = compat
    sys.modules["tensorflow.compat.v1"] = me
    tfp = types.ModuleType("tensorflow_probability")
    tfp.distributions = types.SimpleNamespace(Exponential=_Exponential)
    sys.modules["tensorflow_probability"] = tfp
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    return me
