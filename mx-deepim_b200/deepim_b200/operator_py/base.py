"""Minimal re-statement of the mx.operator.CustomOp / CustomOpProp protocol."""
from __future__ import annotations

import numpy as np
import torch

REGISTRY = {}
_DEFAULT_CTX = {"ctx": None}


def set_default_context(ctx):
    """Context (deepim_b200.context.Context) used by ops created without an explicit one."""
    _DEFAULT_CTX["ctx"] = ctx


def get_context(ctx=None):
    c = ctx or _DEFAULT_CTX["ctx"]
    if c is None:
        raise RuntimeError("deepim_b200.operator_py: no Context set (call set_default_context(Context(...)))")
    return c


def register(name):
    def deco(cls):
        REGISTRY[name] = cls
        cls.op_type = name
        return cls
    return deco


def create(op_type, ctx=None, **str_kwargs):
    """mx.sym.Custom(op_type=...) equivalent: attrs arrive as strings, exactly as in the reference."""
    prop = REGISTRY[op_type](**{k: str(v) for k, v in str_kwargs.items()})
    return prop.create_operator(get_context(ctx), None, None)


def parse_vec(s, n=None):
    """'[a b c]' -> float32 array (np.fromstring(K[1:-1], sep=' ') in the reference)."""
    v = np.array([float(x) for x in s.strip()[1:-1].replace(",", " ").split()], dtype=np.float32)
    if n is not None and v.size != n:
        raise ValueError("expected %d values in %r" % (n, s))
    return v


def parse_bool(s):
    return str(s).lower() == "true"


class CustomOp:
    def assign(self, dst, req, src):
        if req in ("null", None):
            return
        if req in ("write", "inplace"):
            if isinstance(src, (int, float)):
                dst.fill_(src)
            else:
                dst.copy_(src)
        elif req == "add":
            dst.add_(src)
        else:
            raise ValueError("unknown req %r" % (req,))

    def forward(self, is_train, req, in_data, out_data, aux):
        raise NotImplementedError

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        for g, r in zip(in_grad, req):
            self.assign(g, r, 0)


class CustomOpProp:
    def __init__(self, need_top_grad=True):
        self.need_top_grad = need_top_grad

    def infer_type(self, in_type):
        return in_type, [in_type[0]] * len(self.list_outputs()), []

    def list_auxiliary_states(self):
        return []
