"""Host-side mirror of ``TaylorPoly<Interval<F64>>`` -- the number type of the reference's ``--bounds`` mode
(/root/reference/src/interval.rs over /root/reference/src/multivariate_taylor.rs) -- on the device (``gti_*``).

Coefficients are (lo, hi) pairs: arrays carry a trailing axis of length 2.  Every operation runs in the CUDA library
(``csrc/interval_api.cu``); there is no CPU arithmetic here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .taylor import Context, TaylorPanic, _u64, default_context


def _pair(x) -> Tuple[float, float]:
    if isinstance(x, (int, float)):
        return float(x), float(x)
    lo, hi = x
    return float(lo), float(hi)


class IntervalPoly:
    """Device-resident TaylorPoly<Interval<F64>>."""

    __slots__ = ("ctx", "_h")

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    def __del__(self):
        try:
            if self._h and self.ctx.h:
                self.ctx.lib.gti_free(self.ctx.h, self._h)
        except Exception:
            pass

    @classmethod
    def _make(cls, ctx: Context, fn: str, *args) -> "IntervalPoly":
        h = C.c_void_p()
        ctx.check(getattr(ctx.lib, fn)(ctx.h, *args, C.byref(h)))
        return cls(ctx, h)

    def _op(self, fn: str, *args) -> "IntervalPoly":
        return self._make(self.ctx, fn, self._h, *args)

    # -- constructors -----------------------------------------------------------------------
    @classmethod
    def new(cls, coeffs, degrees_p1: Sequence[int], ctx: Optional[Context] = None) -> "IntervalPoly":
        """`coeffs`: shape + (2,) array of (lo, hi) pairs."""
        ctx = ctx or default_context()
        a = np.array(coeffs, dtype=np.float64, order="C")
        assert a.ndim >= 1 and a.shape[-1] == 2, "interval coefficients carry a trailing (lo, hi) axis"
        shape = a.shape[:-1]
        assert len(shape) == len(degrees_p1), "coeffs.ndim() != degrees_p1.len()"
        return cls._make(ctx, "gti_from_host", len(shape), _u64(shape), _u64(degrees_p1), a.ctypes.data_as(C.c_void_p), 1)

    @classmethod
    def from_points(cls, coeffs, degrees_p1: Sequence[int], ctx: Optional[Context] = None) -> "IntervalPoly":
        """f64 coefficients taken as point intervals (Interval::precisely, interval.rs:24-26)."""
        ctx = ctx or default_context()
        a = np.array(coeffs, dtype=np.float64, order="C")
        assert a.ndim == len(degrees_p1), "coeffs.ndim() != degrees_p1.len()"
        return cls._make(ctx, "gti_from_host", a.ndim, _u64(a.shape), _u64(degrees_p1), a.ctypes.data_as(C.c_void_p), 0)

    @classmethod
    def from_scalar(cls, x, ctx: Optional[Context] = None) -> "IntervalPoly":
        lo, hi = _pair(x)
        return cls._make(ctx or default_context(), "gti_from_scalar", lo, hi)

    @classmethod
    def zero_with(cls, degrees_p1: Sequence[int], ctx=None) -> "IntervalPoly":
        return cls._make(ctx or default_context(), "gti_zero_with", len(degrees_p1), _u64(degrees_p1))

    @classmethod
    def var(cls, v: int, x, length: int, ctx=None) -> "IntervalPoly":
        lo, hi = _pair(x)
        return cls._make(ctx or default_context(), "gti_var", v, lo, hi, length)

    @classmethod
    def var_at_zero(cls, v: int, length: int, ctx=None) -> "IntervalPoly":
        return cls._make(ctx or default_context(), "gti_var_at_zero", v, length)

    @classmethod
    def var_with_degrees_p1(cls, v: int, x, degrees_p1: Sequence[int], ctx=None) -> "IntervalPoly":
        lo, hi = _pair(x)
        return cls._make(ctx or default_context(), "gti_var_with_degrees_p1", v, lo, hi, len(degrees_p1), _u64(degrees_p1))

    # -- inspection -------------------------------------------------------------------------
    def num_vars(self) -> int:
        return int(self.ctx.lib.gti_ndim(self._h))

    def array_shape(self) -> tuple:
        n = self.num_vars()
        out = (C.c_uint64 * max(n, 1))()
        self.ctx.lib.gti_shape(self._h, out)
        return tuple(int(out[i]) for i in range(n))

    def degrees_p1(self) -> tuple:
        n = self.num_vars()
        out = (C.c_uint64 * max(n, 1))()
        self.ctx.lib.gti_degrees_p1(self._h, out)
        return tuple(int(out[i]) for i in range(n))

    def array(self) -> np.ndarray:
        a = np.empty(self.array_shape() + (2,), dtype=np.float64)
        self.ctx.check(self.ctx.lib.gti_to_host(self.ctx.h, self._h, a.ctypes.data_as(C.c_void_p)))
        return a

    def constant_term(self) -> Tuple[float, float]:
        out = (C.c_double * 2)()
        self.ctx.check(self.ctx.lib.gti_constant_term(self.ctx.h, self._h, out))
        return out[0], out[1]

    def extract_constant(self) -> Optional[Tuple[float, float]]:
        flag, out = C.c_int(), (C.c_double * 2)()
        self.ctx.check(self.ctx.lib.gti_extract_constant(self.ctx.h, self._h, C.byref(flag), out))
        return (out[0], out[1]) if flag.value else None

    def gather_axis(self, v: int, count: int) -> np.ndarray:
        out = np.empty((count, 2), dtype=np.float64)
        self.ctx.check(self.ctx.lib.gti_gather_axis(self.ctx.h, self._h, v, count, out.ctypes.data_as(C.c_void_p)))
        return out

    # -- operators --------------------------------------------------------------------------
    def _coerce(self, o) -> "IntervalPoly":
        return o if isinstance(o, IntervalPoly) else IntervalPoly.from_scalar(o, self.ctx)

    def __add__(self, o): return self._op("gti_add", self._coerce(o)._h)
    def __sub__(self, o): return self._op("gti_sub", self._coerce(o)._h)
    def __mul__(self, o): return self._op("gti_mul", self._coerce(o)._h)
    def __truediv__(self, o): return self._op("gti_div", self._coerce(o)._h)
    def __neg__(self): return self._op("gti_neg")
    def exp(self): return self._op("gti_exp")
    def log(self): return self._op("gti_log")
    def pow(self, e: int): return self._op("gti_pow", e)
    def derivative(self, v: int, n: int): return self._op("gti_derivative", v, n)
    def taylor_expansion_of_coeff(self, v: int, n: int): return self._op("gti_taylor_expansion_of_coeff", v, n)
    def shift_down(self, v: int, n: int): return self._op("gti_shift_down", v, n)
    def coefficients_of_term(self, v: int, order: int): return self._op("gti_coefficients_of_term", v, order)
    def taylor_polynomial_terms(self, v: int, orders: Sequence[int]):
        return self._op("gti_taylor_polynomial_terms", v, _u64(orders), len(orders))
    def subst_var(self, v: int, subst: "IntervalPoly"): return self._op("gti_subst_var", v, subst._h)
    def truncate_to_degree_p1(self, d: int): return self._op("gti_truncate_to_degree_p1", d)
    def remove_last_variable(self): return self._op("gti_remove_last_variable")
    def extend_to_dim(self, ndim: int, d: int): return self._op("gti_extend_to_dim", ndim, d)

    def __repr__(self) -> str:
        try:
            if not self._h or not self.ctx.h:
                return "IntervalPoly(<released>)"
            return f"IntervalPoly(shape={self.array_shape()}, degrees_p1={self.degrees_p1()})"
        except Exception:
            return "IntervalPoly(<unavailable>)"


class SgclBounds:
    """Enclosures [lo, hi] of the evaluator's direct outputs (gtp_run_sgcl_bounds)."""

    def __init__(self, rest, total, raw_moments, probs):
        self.rest = rest                  # mass of the `rest` generating function
        self.total = total                # Z before the rest mass is added and before clamping to [0, 1]
        self.raw_moments = raw_moments    # E[X] .. E[X^4] (normalised by Z)
        self.probs: List[Tuple[float, float]] = probs   # unnormalised p(0 .. limit-1)


def run_sgcl_bounds(source: str, limit: int = 0, unroll: int = 8, ctx: Optional[Context] = None) -> SgclBounds:
    """The host evaluator over TaylorPoly<Interval<F64>> with all interval arithmetic on the GPU: the enclosure the
    reference's ``--bounds`` mode computes (ratio constants enclosed as Number::from_ratio does; unsimplified GenFun)."""
    ctx = ctx or default_context()
    out = (C.c_double * 12)()
    probs = (C.c_double * max(2 * limit, 2))()
    err = C.create_string_buffer(2048)
    rc = ctx.lib.gtp_run_sgcl_bounds(ctx.h, source.encode(), int(limit), unroll, out, probs, err, 2048)
    if rc != 0:
        raise TaylorPanic(rc, err.value.decode())
    o = list(out)
    return SgclBounds((o[0], o[1]), (o[2], o[3]), [(o[4 + 2 * i], o[5 + 2 * i]) for i in range(4)],
                      [(probs[2 * i], probs[2 * i + 1]) for i in range(limit)])
