"""mx_shim -- a numpy-backed stand-in for the `mxnet` module, just large enough to IMPORT AND RUN the reference's
unmodified custom operators (deepim/operator_py/*.py) and lib/pair_matching/data_pair.py:update_data_batch on the CPU.

Test infrastructure (used by tests/golden/make_golden_mx.py only): the reference's own Python -- bbox extraction, zoom-factor
arithmetic with its mixed float32/float64 scalar promotion, inverse-zoom affines, thresholds, round / +-mean order, the
Transform3D forward / backward, ZoomTrans, GroupPicker, update_data_batch -- executes as written; only the MXNet library
calls underneath it are provided here:

  * NDArray: float32 numpy storage, views on __getitem__, elementwise float32 arithmetic (what MXNet's CPU/GPU kernels do);
  * nd.GridGenerator('affine') + nd.BilinearSampler: THIRD-PARTY semantics (MXNet source is not in /root/reference) --
    implemented from the documented formula in the SAME float32 operation order as oracle/deepim_oracle.c
    (x_t = -1 + j*step, x_s = wx*x_t + tx, x = ((x_s+1)*(W-1))/2, four tap weights formed once, fused multiply-add chain
    emulated in extended precision).  This part therefore stays oracle-defined; everything above it is the reference's code;
  * nd.round = C roundf (half away from zero), which is what MXNet's `round` operator calls;
  * mx.operator.CustomOp / CustomOpProp / register: the protocol of SURVEY 8(b).

numpy-1.x scalar semantics.  The reference (2018, python 2 / numpy 1.x) relies on LEGACY scalar promotion: a float32 numpy
scalar combined with a python int / float gives float64 (numpy >= 2 keeps float32, NEP 50).  `asnumpy()` therefore returns
LegacyArray, whose scalars (LegacyF32) reproduce the numpy-1.x rule, so `tx = zoom_c_x / self.width * 2 - 1` is evaluated in
float64 exactly as it was for the authors.  Nothing else about numpy is altered.
"""
from __future__ import annotations

import sys
import types

import numpy as np


# ------------------------------------------------------------------------------------ numpy-1.x scalars
def _is_py_scalar(x):
    return isinstance(x, (bool, int, float)) and not isinstance(x, np.generic)


class LegacyF32(np.float32):
    """np.float32 with numpy-1.x promotion against python scalars (-> float64); float32 against float32 stays float32."""

    def _bin(self, other, name):
        if _is_py_scalar(other):
            return getattr(np.float64(np.float32(self)), name)(other)
        r = getattr(np.float32, name)(np.float32(self), np.float32(other) if isinstance(other, LegacyF32) else other)
        return LegacyF32(r) if type(r) is np.float32 else r

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        args = [np.float32(x) if isinstance(x, LegacyF32) else x for x in inputs]
        r = getattr(ufunc, method)(*args, **kw)
        return LegacyF32(r) if type(r) is np.float32 else r

    def __neg__(self):
        return LegacyF32(-np.float32(self))

    def __pos__(self):
        return self


for _n in ("add", "sub", "mul", "truediv", "floordiv", "pow", "mod"):
    for _pre in ("__%s__", "__r%s__"):
        _name = _pre % _n
        setattr(LegacyF32, _name, (lambda name: lambda self, other: LegacyF32._bin(self, other, name))(_name))


def _rewrap(r):
    """results of numpy functions on LegacyArray operands keep the legacy scalar behaviour (np.dot / np.sum / np.max ...)"""
    if isinstance(r, np.ndarray):
        if r.ndim == 0:
            r = r[()]
        elif r.dtype == np.float32 and not isinstance(r, LegacyArray):
            return r.view(LegacyArray)
    if type(r) is np.float32:
        return LegacyF32(r)
    if isinstance(r, tuple):
        return tuple(_rewrap(x) for x in r)
    return r


class LegacyArray(np.ndarray):
    """ndarray whose float32 elements come out as LegacyF32 scalars (indexing, iteration, unpacking, reductions)."""

    @staticmethod
    def _wrap(x):
        return LegacyF32(x) if type(x) is np.float32 else x

    def __getitem__(self, k):
        return LegacyArray._wrap(np.ndarray.__getitem__(self, k))

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i]

    def __array_function__(self, func, types, args, kwargs):
        return _rewrap(super().__array_function__(func, types, args, kwargs))

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        args = [np.asarray(x) if isinstance(x, LegacyArray) else (np.float32(x) if isinstance(x, LegacyF32) else x) for x in inputs]
        if "out" in kwargs:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, LegacyArray) else o for o in kwargs["out"])
        return _rewrap(getattr(ufunc, method)(*args, **kwargs))


def _legacy(a):
    a = np.array(a, copy=True)
    return a.view(LegacyArray) if a.dtype == np.float32 else a


# ------------------------------------------------------------------------------------------- NDArray
class Context(object):
    def __init__(self, kind="cpu", idx=0):
        self.device_type, self.device_id = kind, idx

    def __repr__(self):
        return "%s(%d)" % (self.device_type, self.device_id)


_CPU = Context()


def _raw(x):
    """operand of an NDArray kernel: float32 array / float32 scalar (MXNet converts python & numpy scalars to the DType)"""
    if isinstance(x, NDArray):
        return x._a
    if isinstance(x, np.ndarray):
        return x.astype(np.float32)
    return np.float32(x)


class NDArray(object):
    __array_priority__ = 1000.0

    def __init__(self, a, ctx=None):
        self._a = a  # numpy array (possibly a view into a parent NDArray's storage)
        self.context = ctx or _CPU

    # --- structure
    @property
    def shape(self):
        return tuple(self._a.shape)

    @property
    def dtype(self):
        return self._a.dtype.type

    @property
    def size(self):
        return self._a.size

    def asnumpy(self):
        return _legacy(self._a)

    def asscalar(self):
        assert self._a.size == 1
        return LegacyArray._wrap(self._a.reshape(-1)[0])

    def reshape(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return NDArray(self._a.reshape(shape), self.context)

    def copy(self):
        return NDArray(self._a.copy(), self.context)

    def as_in_context(self, ctx):
        return self

    def astype(self, dt):
        return NDArray(self._a.astype(dt), self.context)

    def __len__(self):
        return self._a.shape[0]

    def __getitem__(self, k):
        r = self._a[k]
        if not isinstance(r, np.ndarray):  # MXNet returns a 1-element NDArray for a full index
            kk = k if isinstance(k, tuple) else (k,)
            r = self._a[kk[:-1] + (slice(kk[-1], kk[-1] + 1),)] if isinstance(kk[-1], (int, np.integer)) else np.asarray(r)
        return NDArray(r, self.context)

    def __setitem__(self, k, v):
        self._a[k] = _raw(v).reshape(np.shape(self._a[k])) if isinstance(v, NDArray) and v._a.size == np.size(self._a[k]) \
            else _raw(v)

    # --- float32 elementwise arithmetic
    def _op(self, other, f, rev=False):
        a, b = self._a, _raw(other)
        r = f(b, a) if rev else f(a, b)
        return NDArray(np.asarray(r, dtype=self._a.dtype if self._a.dtype != np.float64 else np.float32), self.context)

    def __add__(self, o): return self._op(o, np.add)
    def __radd__(self, o): return self._op(o, np.add, True)
    def __sub__(self, o): return self._op(o, np.subtract)
    def __rsub__(self, o): return self._op(o, np.subtract, True)
    def __mul__(self, o): return self._op(o, np.multiply)
    def __rmul__(self, o): return self._op(o, np.multiply, True)
    def __truediv__(self, o): return self._op(o, np.divide)
    def __rtruediv__(self, o): return self._op(o, np.divide, True)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __neg__(self): return NDArray(-self._a, self.context)

    def _iop(self, o, f):
        self._a[...] = f(self._a, _raw(o))
        return self

    def __iadd__(self, o): return self._iop(o, np.add)
    def __isub__(self, o): return self._iop(o, np.subtract)
    def __imul__(self, o): return self._iop(o, np.multiply)
    def __itruediv__(self, o): return self._iop(o, np.divide)
    __idiv__ = __itruediv__

    def __repr__(self):
        return "<shim NDArray %s>\n%s" % (self.shape, self._a)


# ------------------------------------------------------------------------------------------ nd functions
def _dtype(dtype):
    return np.dtype(dtype if dtype is not None else np.float32)


def nd_array(src, ctx=None, dtype=None):
    if isinstance(src, NDArray):
        src = src._a
    return NDArray(np.array(src, dtype=_dtype(dtype)), ctx)  # default dtype float32 (mx.nd.array of a float64 numpy array casts)


def nd_zeros(shape, ctx=None, dtype=None, **kw):
    return NDArray(np.zeros(shape, _dtype(dtype)), ctx)


def nd_ones(shape, ctx=None, dtype=None, **kw):
    return NDArray(np.ones(shape, _dtype(dtype)), ctx)


def nd_zeros_like(a, ctx=None, dtype=None, **kw):
    return NDArray(np.zeros(a.shape, _dtype(dtype)), ctx)


def nd_round(a):
    x = a._a
    return NDArray((np.sign(x) * np.floor(np.abs(x) + np.float32(0.5))).astype(np.float32), a.context)  # roundf


def _fma32(x, y, z):
    """fmaf(x, y, z) for float32 arrays: the product is exact in extended precision, one rounding on the way back"""
    return (x.astype(np.longdouble) * y.astype(np.longdouble) + z.astype(np.longdouble)).astype(np.float32)


def nd_GridGenerator(data, transform_type="affine", target_shape=None):
    assert transform_type == "affine"
    H, W = int(target_shape[0]), int(target_shape[1])
    aff = data._a.reshape(-1, 6).astype(np.float32)
    stepx, stepy = np.float32(2.0 / (W - 1)), np.float32(2.0 / (H - 1))
    xt = np.float32(-1.0) + np.arange(W, dtype=np.float32) * stepx
    yt = np.float32(-1.0) + np.arange(H, dtype=np.float32) * stepy
    out = np.zeros((aff.shape[0], 2, H, W), np.float32)
    for b in range(aff.shape[0]):
        wx, sx, tx, sy, wy, ty = aff[b]
        assert sx == 0 and sy == 0, "only the axis-aligned affines the reference builds are supported"
        out[b, 0] = (wx * xt + tx)[None, :]
        out[b, 1] = (wy * yt + ty)[:, None]
    return NDArray(out, data.context)


def nd_BilinearSampler(data, grid):
    src, g = data._a.astype(np.float32), grid._a
    B, C, H, W = src.shape
    out = np.zeros((B, C) + g.shape[2:], np.float32)
    one, two = np.float32(1.0), np.float32(2.0)
    for b in range(B):
        xr = ((g[b, 0] + one) * np.float32(W - 1)) / two
        yr = ((g[b, 1] + one) * np.float32(H - 1)) / two
        fx0, fy0 = np.floor(xr), np.floor(yr)
        wx1, wy1 = one - (xr - fx0), one - (yr - fy0)
        x0 = np.clip(fx0, -4, W + 4).astype(np.int64)
        y0 = np.clip(fy0, -4, H + 4).astype(np.int64)
        ax, ay = one - wx1, one - wy1
        wa, wb, wc, wd = wy1 * wx1, wy1 * ax, ay * wx1, ay * ax

        def tap(img, yy, xx):
            ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
            return np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], np.float32(0.0)).astype(np.float32)

        for c in range(C):
            img = src[b, c]
            tl, tr, bl, br = tap(img, y0, x0), tap(img, y0, x0 + 1), tap(img, y0 + 1, x0), tap(img, y0 + 1, x0 + 1)
            out[b, c] = _fma32(br, wd, _fma32(bl, wc, _fma32(tr, wb, tl * wa)))
    return NDArray(out, data.context)


def nd_batch_dot(a, b, transpose_a=False, transpose_b=False):
    x, y = a._a.astype(np.float32), b._a.astype(np.float32)
    if transpose_a:
        x = x.transpose(0, 2, 1)
    if transpose_b:
        y = y.transpose(0, 2, 1)
    return NDArray(np.matmul(x, y).astype(np.float32), a.context)


def nd_dot(a, b):
    return NDArray(np.dot(a._a, b._a).astype(np.float32), a.context)


def _un(f):
    return lambda a, **kw: NDArray(f(a._a).astype(np.float32), a.context)


def _bin(f):
    return lambda a, b: NDArray(np.asarray(f(_raw(a), _raw(b)), np.float32), a.context if isinstance(a, NDArray) else b.context)


def nd_sum(a, axis=None, keepdims=False):
    return NDArray(np.asarray(np.sum(a._a, axis=axis, keepdims=keepdims, dtype=np.float32)), a.context)


def nd_slice_axis(a, axis, begin, end):
    idx = [slice(None)] * a._a.ndim
    idx[axis] = slice(begin, end)
    return NDArray(a._a[tuple(idx)].copy(), a.context)


def nd_split(a, axis, num_outputs):
    return [NDArray(p.copy(), a.context) for p in np.split(a._a, num_outputs, axis=axis)]


def nd_concat(*arrs, **kw):
    return NDArray(np.concatenate([x._a for x in arrs], axis=kw.get("dim", 1)), arrs[0].context)


_ND = {
    "NDArray": NDArray, "array": nd_array, "zeros": nd_zeros, "ones": nd_ones, "zeros_like": nd_zeros_like, "round": nd_round,
    "GridGenerator": nd_GridGenerator, "BilinearSampler": nd_BilinearSampler, "batch_dot": nd_batch_dot, "dot": nd_dot,
    "exp": _un(np.exp), "sqrt": _un(np.sqrt), "abs": _un(np.abs), "sum": nd_sum,
    "add": _bin(np.add), "subtract": _bin(np.subtract), "multiply": _bin(np.multiply), "divide": _bin(np.divide),
    "maximum": _bin(np.maximum), "minimum": _bin(np.minimum), "broadcast_mul": _bin(np.multiply),
    "broadcast_add": _bin(np.add), "broadcast_sub": _bin(np.subtract), "broadcast_div": _bin(np.divide),
    "transpose": lambda a, axes=None: NDArray(np.transpose(a._a, axes).copy(), a.context),
    "expand_dims": lambda a, axis: NDArray(np.expand_dims(a._a, axis), a.context),
    "tile": lambda a, reps: NDArray(np.tile(a._a, reps), a.context),
    "slice_axis": nd_slice_axis, "split": nd_split, "concat": nd_concat,
}


# ----------------------------------------------------------------------------------- operator protocol
_REGISTRY = {}


class CustomOp(object):
    def assign(self, dst, req, src):
        """req in {null, write, inplace, add}"""
        if req == "null":
            return
        val = _raw(src)
        if req in ("write", "inplace"):
            dst._a[...] = val
        elif req == "add":
            dst._a[...] += val
        else:
            raise ValueError("unknown req %r" % (req,))


class CustomOpProp(object):
    def __init__(self, need_top_grad=False):
        self.need_top_grad_ = need_top_grad

    def infer_type(self, in_type):
        return in_type, [in_type[0]] * len(self.list_outputs()), []


def register(name):
    def deco(cls):
        _REGISTRY[name] = cls
        return cls
    return deco


def run_op(op_type, inputs, is_train=False, **attrs):
    """What mx.sym.Custom(op_type=..., **tensors, **string attrs) + one executor forward does: build the Prop from the string
    attrs, infer the output shapes, create the operator, call forward with req='write'.  Returns (outputs as numpy, op, io)."""
    prop = _REGISTRY[op_type](**{k: str(v) for k, v in attrs.items()})
    in_nd = [nd_array(np.asarray(x, np.float32)) for x in inputs]
    _, out_shapes, _ = prop.infer_shape([list(x.shape) for x in in_nd])
    out_nd = [nd_zeros(tuple(int(d) for d in s)) for s in out_shapes]
    op = prop.create_operator(_CPU, [x.shape for x in in_nd], [np.float32] * len(in_nd))
    op.forward(is_train, ["write"] * len(out_nd), in_nd, out_nd, [])
    return [np.array(o._a) for o in out_nd], op, (prop, in_nd, out_nd)


def run_op_backward(op, io, out_grads):
    prop, in_nd, out_nd = io
    in_grad = [nd_zeros(x.shape) for x in in_nd]
    og = [nd_array(np.asarray(g, np.float32)) for g in out_grads]
    op.backward(["write"] * len(in_nd), og, in_nd, out_nd, in_grad, [])
    return [np.array(g._a) for g in in_grad]


# ------------------------------------------------------------------------------------------ installation
def install():
    """Put the stand-in modules into sys.modules (idempotent).  Must run BEFORE the reference files are imported."""
    if "mxnet" in sys.modules and getattr(sys.modules["mxnet"], "__shim__", False):
        return sys.modules["mxnet"]
    mx = types.ModuleType("mxnet")
    mx.__shim__ = True
    nd = types.ModuleType("mxnet.ndarray")
    for k, v in _ND.items():
        setattr(nd, k, v)
    op = types.ModuleType("mxnet.operator")
    op.CustomOp, op.CustomOpProp, op.register = CustomOp, CustomOpProp, register
    sym = types.ModuleType("mxnet.symbol")  # only referenced by the __main__ demo blocks of the reference files
    mx.nd = mx.ndarray = nd
    mx.operator, mx.sym, mx.symbol = op, sym, sym
    mx.cpu = lambda i=0: _CPU
    mx.gpu = lambda i=0: _CPU
    mx.Context = Context
    sys.modules.update({"mxnet": mx, "mxnet.ndarray": nd, "mxnet.nd": nd, "mxnet.operator": op, "mxnet.symbol": sym})
    if "distutils.util" not in sys.modules:  # python >= 3.12 has no distutils; transform3d.py imports strtobool from it
        try:
            import distutils.util  # noqa: F401
        except Exception:
            du, dutil = types.ModuleType("distutils"), types.ModuleType("distutils.util")

            def strtobool(v):
                v = str(v).lower()
                if v in ("y", "yes", "t", "true", "on", "1"):
                    return 1
                if v in ("n", "no", "f", "false", "off", "0"):
                    return 0
                raise ValueError("invalid truth value %r" % (v,))
            dutil.strtobool = strtobool
            du.util = dutil
            sys.modules["distutils"], sys.modules["distutils.util"] = du, dutil
    # numpy-2 removed aliases that lib/pair_matching/RT_transform.py touches at import time (same 3-line shim as make_golden.py)
    for name, val in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, val)
    if not hasattr(np, "maximum_sctype"):
        np.maximum_sctype = lambda t: np.float64
    return mx
