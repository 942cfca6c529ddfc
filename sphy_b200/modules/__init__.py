"""The reference's process modules, same names and call signatures, operating on device maps.

The reference passes the PCRaster namespace `pcr` into every process function (`f(self?, pcr, maps...) -> maps`,
e.g. /root/reference/rootzone.py:63, modules/routing.py:25).  These modules keep those signatures: `pcr` is accepted
and ignored (there is no raster library behind it any more), map arguments are float32 CUDA tensors over the active
cells (or Python floats for non-spatial values), results are new tensors.  Each function is one launch of
`sphy_module_op` -- the same __device__ function the fused kernel `sphy_wb_step` composes -- or of the routing /
glacier / lake-reservoir entry points.  `SphyModel.dynamic()` itself uses the fused kernel; these modules are the
plug-in surface for code written against the reference's modules, and what tests/test_gpu_modules.py checks
function by function against the oracle.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib

f32 = np.float32


class _Runtime:
    lib = None
    ctx = None
    device = None
    n = 0


RT = _Runtime()


def bind(model):
    """Attach the modules to a model's context (done by SphyModel.__init__; the last model bound wins)."""
    RT.lib, RT.ctx, RT.device, RT.n = model.lib, model._ctx, model.device, model.n


def _stream():
    return C.c_void_p(torch.cuda.current_stream(RT.device).cuda_stream)


def as_param(x, keep):
    """tensor | ndarray | number -> SphyParam"""
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float32 or not x.is_cuda:
            x = x.to(device=RT.device, dtype=torch.float32)
        x = x.contiguous()
        keep.append(x)
        return _lib.SphyParam(x.data_ptr(), 0.0)
    if isinstance(x, np.ndarray):
        if x.ndim == 0:
            return _lib.SphyParam(None, float(f32(x)))
        t = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(RT.device)
        keep.append(t)
        return _lib.SphyParam(t.data_ptr(), 0.0)
    return _lib.SphyParam(None, float(f32(x)))


def op(name, inputs, n_out=1, k0=0.0, doy=1):
    """Launch one per-cell process function over the whole map."""
    if RT.lib is None:
        raise RuntimeError("sphy_b200.modules is not bound to a model (no CPU fallback exists)")
    keep = []
    params = (_lib.SphyParam * len(inputs))(*[as_param(x, keep) for x in inputs])
    outs = [torch.empty(RT.n, dtype=torch.float32, device=RT.device) for _ in range(n_out)]
    outp = (C.c_void_p * n_out)(*[t.data_ptr() for t in outs])
    _lib.check(RT.lib, RT.ctx, RT.lib.sphy_module_op(RT.ctx, _lib.OP[name], params, len(inputs), outp, n_out, RT.n,
                                                     C.c_float(k0), int(doy), _stream()), f"sphy_module_op({name})")
    return outs[0] if n_out == 1 else tuple(outs)


def exp_neg_inv(x):
    """REAL4(exp(-1/x)): `-1/x` is a REAL4 division for a map, a Python-double division for a cfg float."""
    if isinstance(x, torch.Tensor):
        out = torch.empty_like(x)
        _lib.check(RT.lib, RT.ctx, RT.lib.sphy_map_unary(RT.ctx, _lib.UNARY_EXP_NEG_INV, x.data_ptr(), out.data_ptr(), x.numel(),
                                                         _stream()), "sphy_map_unary")
        return out
    if isinstance(x, np.ndarray):
        with np.errstate(all="ignore"):
            return np.exp((f32(-1) / x.astype(np.float32)).astype(np.float64)).astype(np.float32)
    return float(f32(math.exp(float(f32(-1 / x)))))


def exp_neg(x):
    """REAL4(exp(-x))"""
    if isinstance(x, torch.Tensor):
        out = torch.empty_like(x)
        _lib.check(RT.lib, RT.ctx, RT.lib.sphy_map_unary(RT.ctx, _lib.UNARY_EXP_NEG, x.data_ptr(), out.data_ptr(), x.numel(), _stream()),
                   "sphy_map_unary")
        return out
    if isinstance(x, np.ndarray):
        return np.exp((-x.astype(np.float32)).astype(np.float64)).astype(np.float32)
    return float(f32(math.exp(float(f32(-x)))))


from . import (ET, advanced_routing, glacier, groundwater, hargreaves, lakes, mmf, musle, pedotransfer, reservoirs, rootzone, routing,  # noqa: E402,F401
               sediment_transport, snow, subzone)
