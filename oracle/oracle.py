"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to ``liboracle.so`` (the C++ restatement of the reference's
``TaylorPoly<T>`` / ``TaylorExpansion<T>``, see ``taylor_oracle.hpp``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module; the product path (``genfer_b200``) never does.

The classes mirror the reference's operator surface
(/root/reference/src/multivariate_taylor.rs, src/univariate_taylor.rs) so that the parity
tests read like the reference's own unit tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
UMAX = 2**64 - 1  # usize::MAX sentinel ("unbounded degree")


def _cpu_stamp() -> str:
    """Identity of the host CPU the library was compiled for (-march=native): model name + ISA flags."""
    import hashlib
    model, flags = "", ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name") and not model:
                    model = ln.split(":", 1)[1].strip()
                elif ln.startswith("flags") and not flags:
                    flags = ln.split(":", 1)[1].strip()
                if model and flags:
                    break
    except OSError:
        pass
    return model + " | " + hashlib.sha1(flags.encode()).hexdigest()[:16]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ -O3 -march=native -ffp-contract=off, BASELINE.md's
    CPU-baseline recipe).  -march=native ties the binary to the build host, so a stamp records the CPU it was built
    on and a different host (the GPU box) rebuilds once (~30 s) instead of running foreign code."""
    import fcntl
    import glob
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle_eval.cpp", "taylor_oracle.hpp", "Makefile")]
    srcs += glob.glob(os.path.join(_HERE, "..", "genfer_b200", "csrc", "evaluator", "*.hpp"))
    stamp_path = os.path.join(_HERE, "liboracle.stamp")
    stamp = _cpu_stamp()

    def fresh() -> bool:
        if not os.path.exists(_LIB_PATH) or not os.path.exists(stamp_path):
            return False
        if open(stamp_path).read().strip() != stamp:
            return False
        return all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)

    if not force and fresh():
        return _LIB_PATH
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not fresh():
            subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
            with open(stamp_path, "w") as f:
                f.write(stamp + "\n")
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def _declare(L: C.CDLL) -> None:
    L.orc_last_error.restype = C.c_char_p
    for P in ("orc_f64_", "orc_iv_"):
        def f(name, restype, *argtypes):
            fn = getattr(L, P + name)
            fn.restype = restype
            fn.argtypes = list(argtypes)
        vp = C.c_void_p
        f("new", vp, C.c_int, _u64p, _u64p, _f64p)
        f("free", None, vp)
        f("ndim", C.c_int, vp)
        f("len", C.c_uint64, vp)
        f("shape", None, vp, _u64p)
        f("degrees", None, vp, _u64p)
        f("data", None, vp, _f64p)
        f("eq", C.c_int, vp, vp)
        for b in ("add", "sub", "mul", "div"):
            f(b, vp, vp, vp)
        for u in ("neg", "exp", "log", "remove_last_variable"):
            f(u, vp, vp)
        f("pow", vp, vp, C.c_uint32)
        for g in ("derivative", "taylor_expansion_of_coeff", "shift_down", "coefficients_of_term",
                  "taylor_polynomial"):
            f(g, vp, vp, C.c_uint64, C.c_uint64)
        f("subst_var", vp, vp, C.c_uint64, vp)
        f("taylor_polynomial_terms", vp, vp, C.c_uint64, _u64p, C.c_int)
        f("truncate_to_degree_p1", vp, vp, C.c_uint64)
        f("extend_to_dim", vp, vp, C.c_uint64, C.c_uint64)
        f("extend", vp, vp, _u64p, C.c_int)
        f("zero_with", vp, _u64p, C.c_int)
        f("var", vp, C.c_uint64, _f64p, C.c_uint64)
        f("var_at_zero", vp, C.c_uint64, C.c_uint64)
        f("var_with_degrees_p1", vp, C.c_uint64, _f64p, _u64p, C.c_int)
        f("coefficient", C.c_int, vp, _u64p, C.c_int, _f64p)
        f("constant_term", None, vp, _f64p)
        f("evaluate_all_one", None, vp, _f64p)
        f("is_zero", C.c_int, vp)
        f("is_one", C.c_int, vp)
        f("extract_linear", C.c_int, vp, _f64p, _f64p, _u64p)
        f("te_const", vp, _f64p)
        f("te_poly", vp, _f64p, C.c_uint64)
        f("te_var", vp, _f64p, C.c_uint64)
        f("te_free", None, vp)
        f("te_is_const", C.c_int, vp)
        f("te_len", C.c_uint64, vp)
        f("te_data", None, vp, _f64p)
        f("te_eq", C.c_int, vp, vp)
        for b in ("te_add", "te_sub", "te_mul", "te_div", "te_subst"):
            f(b, vp, vp, vp)
        for u in ("te_neg", "te_exp", "te_log"):
            f(u, vp, vp)
        f("te_pow", vp, vp, C.c_uint32)
        f("te_taylor_expansion_of_coeff", vp, vp, C.c_uint64)
        f("te_coeff", C.c_int, vp, C.c_uint64, _f64p)
        f("te_derivative", C.c_int, vp, C.c_uint64, _f64p)
    L.orc_mul_raw.restype = None
    L.orc_mul_raw.argtypes = [C.c_int, _u64p, _f64p, _u64p, _f64p, _u64p, _f64p]
    L.orc_mul_rows.restype = C.c_double
    L.orc_mul_rows.argtypes = [C.c_int, _u64p, _f64p, _u64p, _f64p, _u64p, _f64p, _u64p, C.c_int]
    L.orc_mul_macs.restype = C.c_double
    L.orc_mul_macs.argtypes = [C.c_int, _u64p, _u64p, _u64p]
    L.orc_next_up.restype = C.c_double
    L.orc_next_up.argtypes = [C.c_double]
    L.orc_next_down.restype = C.c_double
    L.orc_next_down.argtypes = [C.c_double]


class OracleError(RuntimeError):
    """The reference would panic here (assert!/index out of bounds)."""


def _u64(xs: Iterable[int]):
    xs = [int(x) for x in xs]
    return (C.c_uint64 * max(len(xs), 1))(*xs)


def _f64(a: np.ndarray):
    return a.ctypes.data_as(_f64p)


class TaylorPoly:
    """Oracle TaylorPoly (kind 'f64' -> T = F64, kind 'iv' -> T = Interval<F64>)."""

    __slots__ = ("_h", "kind")

    def __init__(self, handle, kind: str = "f64"):
        self._h = handle
        self.kind = kind
        if not handle:
            raise OracleError(lib().orc_last_error().decode())

    # -- plumbing ---------------------------------------------------------------------
    @staticmethod
    def _fn(kind: str, name: str):
        return getattr(lib(), ("orc_f64_" if kind == "f64" else "orc_iv_") + name)

    def _f(self, name: str):
        return self._fn(self.kind, name)

    def __del__(self):
        try:
            if self._h:
                self._f("free")(self._h)
        except Exception:
            pass

    @property
    def _w(self) -> int:
        return 1 if self.kind == "f64" else 2

    # -- constructors (multivariate_taylor.rs:33-46, 208-259, 626-656) -----------------
    @classmethod
    def new(cls, coeffs, degrees_p1: Sequence[int], kind: str = "f64") -> "TaylorPoly":
        a = np.array(coeffs, dtype=np.float64, order="C")  # (ascontiguousarray would promote 0-d to 1-d)
        shape = a.shape if kind == "f64" else a.shape[:-1]
        assert len(shape) == len(degrees_p1), "coeffs.ndim() != degrees_p1.len()"
        return cls(cls._fn(kind, "new")(len(shape), _u64(shape), _u64(degrees_p1), _f64(a)), kind)

    @classmethod
    def from_coeffs(cls, coeffs, kind: str = "f64") -> "TaylorPoly":
        a = np.asarray(coeffs, dtype=np.float64)
        shape = a.shape if kind == "f64" else a.shape[:-1]
        return cls.new(a, shape, kind)

    @classmethod
    def from_scalar(cls, x, kind: str = "f64") -> "TaylorPoly":
        a = np.array(x, dtype=np.float64)
        return cls.new(a, (), kind)

    @classmethod
    def zero(cls, kind: str = "f64") -> "TaylorPoly":
        return cls.from_scalar(0.0 if kind == "f64" else [0.0, 0.0], kind)

    @classmethod
    def one(cls, kind: str = "f64") -> "TaylorPoly":
        return cls.from_scalar(1.0 if kind == "f64" else [1.0, 1.0], kind)

    @classmethod
    def from_u32(cls, c: int, kind: str = "f64") -> "TaylorPoly":
        return cls.from_scalar(float(c) if kind == "f64" else [float(c), float(c)], kind)

    @classmethod
    def zero_with(cls, degrees_p1: Sequence[int], kind: str = "f64") -> "TaylorPoly":
        return cls(cls._fn(kind, "zero_with")(_u64(degrees_p1), len(degrees_p1)), kind)

    @classmethod
    def var(cls, v: int, x, length: int, kind: str = "f64") -> "TaylorPoly":
        xa = np.atleast_1d(np.array(x, dtype=np.float64))
        return cls(cls._fn(kind, "var")(v, _f64(xa), length), kind)

    @classmethod
    def var_at_zero(cls, v: int, length: int, kind: str = "f64") -> "TaylorPoly":
        return cls(cls._fn(kind, "var_at_zero")(v, length), kind)

    @classmethod
    def var_with_degrees_p1(cls, v: int, x, degrees_p1: Sequence[int], kind: str = "f64") -> "TaylorPoly":
        xa = np.atleast_1d(np.array(x, dtype=np.float64))
        return cls(cls._fn(kind, "var_with_degrees_p1")(v, _f64(xa), _u64(degrees_p1), len(degrees_p1)), kind)

    # -- accessors --------------------------------------------------------------------
    def num_vars(self) -> int:
        return self._f("ndim")(self._h)

    def array_shape(self) -> tuple:
        n = self.num_vars()
        out = (C.c_uint64 * max(n, 1))()
        self._f("shape")(self._h, out)
        return tuple(int(out[i]) for i in range(n))

    def shape(self) -> tuple:
        """Reference `shape()` returns degrees_p1 (multivariate_taylor.rs:53-56)."""
        n = self.num_vars()
        out = (C.c_uint64 * max(n, 1))()
        self._f("degrees")(self._h, out)
        return tuple(int(out[i]) for i in range(n))

    degrees_p1 = property(lambda self: self.shape())

    def array(self) -> np.ndarray:
        shp = self.array_shape()
        full = shp if self.kind == "f64" else shp + (2,)
        a = np.empty(full, dtype=np.float64)
        self._f("data")(self._h, _f64(a))
        return a

    def is_zero(self) -> bool:
        return bool(self._f("is_zero")(self._h))

    def is_one(self) -> bool:
        return bool(self._f("is_one")(self._h))

    def is_constant(self) -> bool:
        return int(self._f("len")(self._h)) == 1

    def constant_term(self):
        out = np.empty(self._w)
        self._f("constant_term")(self._h, _f64(out))
        return float(out[0]) if self.kind == "f64" else out

    def evaluate_all_one(self):
        out = np.empty(self._w)
        self._f("evaluate_all_one")(self._h, _f64(out))
        return float(out[0]) if self.kind == "f64" else out

    def coefficient(self, index: Sequence[int]):
        out = np.empty(self._w)
        if self._f("coefficient")(self._h, _u64(index), len(index), _f64(out)):
            raise OracleError(lib().orc_last_error().decode())
        return float(out[0]) if self.kind == "f64" else out

    def extract_linear(self) -> Optional[tuple]:
        c, m = np.empty(self._w), np.empty(self._w)
        v = C.c_uint64()
        if not self._f("extract_linear")(self._h, _f64(c), _f64(m), C.byref(v)):
            return None
        if self.kind == "f64":
            return float(c[0]), float(m[0]), int(v.value)
        return c, m, int(v.value)

    # -- operators --------------------------------------------------------------------
    def _bin(self, name: str, other: "TaylorPoly") -> "TaylorPoly":
        return TaylorPoly(self._f(name)(self._h, other._h), self.kind)

    def _un(self, name: str, *args) -> "TaylorPoly":
        return TaylorPoly(self._f(name)(self._h, *args), self.kind)

    __add__ = lambda self, o: self._bin("add", o)
    __sub__ = lambda self, o: self._bin("sub", o)
    __mul__ = lambda self, o: self._bin("mul", o)
    __truediv__ = lambda self, o: self._bin("div", o)
    __neg__ = lambda self: self._un("neg")

    def __eq__(self, other) -> bool:  # derive(PartialEq): stored shape, data and degrees
        return bool(self._f("eq")(self._h, other._h))

    def __ne__(self, other) -> bool:
        return not self.__eq__(other)

    __hash__ = None

    def exp(self): return self._un("exp")
    def log(self): return self._un("log")
    def pow(self, e: int): return self._un("pow", e)
    def derivative(self, v: int, n: int): return self._un("derivative", v, n)
    def taylor_expansion_of_coeff(self, v: int, n: int): return self._un("taylor_expansion_of_coeff", v, n)
    def shift_down(self, v: int, n: int): return self._un("shift_down", v, n)
    def coefficients_of_term(self, v: int, order: int): return self._un("coefficients_of_term", v, order)
    def taylor_polynomial(self, v: int, order: int): return self._un("taylor_polynomial", v, order)
    def taylor_polynomial_terms(self, v: int, orders: Sequence[int]):
        return self._un("taylor_polynomial_terms", v, _u64(orders), len(orders))
    def subst_var(self, v: int, subst: "TaylorPoly"):
        return TaylorPoly(self._f("subst_var")(self._h, v, subst._h), self.kind)
    def truncate_to_degree_p1(self, d: int): return self._un("truncate_to_degree_p1", d)
    def remove_last_variable(self): return self._un("remove_last_variable")
    def extend_to_dim(self, ndim: int, d: int): return self._un("extend_to_dim", ndim, d)
    def extend(self, new_size: Sequence[int]): return self._un("extend", _u64(new_size), len(new_size))

    def __repr__(self) -> str:
        if not self._h:      # a constructor that raised: pytest reprs the half-built object while formatting the traceback
            return "TaylorPoly(<invalid>)"
        return f"TaylorPoly({list(self.shape())}, {self.array().tolist()})"


def taylor(coeffs, degrees_p1: Optional[Sequence[int]] = None, kind: str = "f64") -> TaylorPoly:
    """The reference's `taylor!` test macro (multivariate_taylor.rs:659-692)."""
    if degrees_p1 is None:
        return TaylorPoly.from_coeffs(coeffs, kind)
    return TaylorPoly.new(coeffs, degrees_p1, kind)


class TaylorExpansion:
    """Oracle univariate TaylorExpansion<T> (univariate_taylor.rs:9-13)."""

    __slots__ = ("_h", "kind")

    def __init__(self, handle, kind: str = "f64"):
        self._h = handle
        self.kind = kind
        if not handle:
            raise OracleError(lib().orc_last_error().decode())

    def _f(self, name):
        return TaylorPoly._fn(self.kind, name)

    def __del__(self):
        try:
            if self._h:
                self._f("te_free")(self._h)
        except Exception:
            pass

    @classmethod
    def constant(cls, x, kind="f64"):
        xa = np.atleast_1d(np.array(x, dtype=np.float64))
        return cls(TaylorPoly._fn(kind, "te_const")(_f64(xa)), kind)

    @classmethod
    def from_coefficients(cls, xs, kind="f64"):
        a = np.ascontiguousarray(xs, dtype=np.float64)
        n = a.shape[0]
        return cls(TaylorPoly._fn(kind, "te_poly")(_f64(a), n), kind)

    @classmethod
    def var(cls, x, order: int, kind="f64"):
        xa = np.atleast_1d(np.array(x, dtype=np.float64))
        return cls(TaylorPoly._fn(kind, "te_var")(_f64(xa), order), kind)

    @classmethod
    def zero(cls, kind="f64"): return cls.constant(0.0 if kind == "f64" else [0.0, 0.0], kind)
    @classmethod
    def one(cls, kind="f64"): return cls.constant(1.0 if kind == "f64" else [1.0, 1.0], kind)

    def is_const(self) -> bool:
        return bool(self._f("te_is_const")(self._h))

    def coeffs(self) -> np.ndarray:
        n = int(self._f("te_len")(self._h))
        w = 1 if self.kind == "f64" else 2
        a = np.empty((n,) if w == 1 else (n, 2))
        self._f("te_data")(self._h, _f64(a))
        return a

    def order(self) -> int:
        return UMAX if self.is_const() else int(self._f("te_len")(self._h))

    def coeff(self, order: int):
        out = np.empty(1 if self.kind == "f64" else 2)
        if self._f("te_coeff")(self._h, order, _f64(out)):
            raise OracleError(lib().orc_last_error().decode())
        return float(out[0]) if self.kind == "f64" else out

    def derivative(self, order: int):
        out = np.empty(1 if self.kind == "f64" else 2)
        if self._f("te_derivative")(self._h, order, _f64(out)):
            raise OracleError(lib().orc_last_error().decode())
        return float(out[0]) if self.kind == "f64" else out

    def _bin(self, name, o): return TaylorExpansion(self._f(name)(self._h, o._h), self.kind)
    def _un(self, name, *a): return TaylorExpansion(self._f(name)(self._h, *a), self.kind)
    __add__ = lambda s, o: s._bin("te_add", o)
    __sub__ = lambda s, o: s._bin("te_sub", o)
    __mul__ = lambda s, o: s._bin("te_mul", o)
    __truediv__ = lambda s, o: s._bin("te_div", o)
    __neg__ = lambda s: s._un("te_neg")
    def __eq__(self, o): return bool(self._f("te_eq")(self._h, o._h))
    def __ne__(self, o): return not self.__eq__(o)
    __hash__ = None
    def exp(self): return self._un("te_exp")
    def log(self): return self._un("te_log")
    def pow(self, e: int): return self._un("te_pow", e)
    def subst(self, s): return self._bin("te_subst", s)
    def taylor_expansion_of_coeff(self, n: int): return self._un("te_taylor_expansion_of_coeff", n)

    def __repr__(self):
        if not self._h:
            return "TaylorExpansion(<invalid>)"
        return f"TaylorExpansion({'Constant' if self.is_const() else 'Polynomial'}, {self.coeffs().tolist()})"


# -- raw helpers for parity tests / the bench CPU leg -------------------------------------
def mul_raw(x: np.ndarray, y: np.ndarray, rshape: Sequence[int]) -> np.ndarray:
    """General product (multivariate_taylor.rs:984-1012) on contiguous f64 arrays."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    r = np.zeros(tuple(rshape), dtype=np.float64)
    lib().orc_mul_raw(x.ndim, _u64(x.shape), _f64(x), _u64(y.shape), _f64(y), _u64(r.shape), _f64(r))
    return r


def mul_rows(x: np.ndarray, y: np.ndarray, rshape: Sequence[int], rows: Sequence[int]):
    """Bounded sample: only leading-axis output rows `rows`.  Returns (result, MACs executed)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    r = np.zeros(tuple(rshape), dtype=np.float64)
    macs = lib().orc_mul_rows(x.ndim, _u64(x.shape), _f64(x), _u64(y.shape), _f64(y), _u64(r.shape), _f64(r),
                              _u64(rows), len(rows))
    return r, macs


class SgclResult:
    """Outcome of running an SGCL program end to end (report = the reference's stdout with --no-timing)."""

    def __init__(self, report, moments, probs, normalized_probs, moment_bounds=None, prob_bounds=None):
        self.moment_bounds = moment_bounds    # 11 (lo, hi) pairs behind `moments` (--bounds runs: the result proper)
        self.prob_bounds = prob_bounds        # (lo, hi) of the unnormalised p(i)
        self.report = report
        (self.total, self.mean, self.raw2, self.raw3, self.raw4, self.stddev, self.variance, self.central3,
         self.central4, self.skewness, self.kurtosis) = moments
        self.moments = list(moments)
        self.probs = probs
        self.normalized_probs = normalized_probs


def run_sgcl(source: str, limit: Optional[int] = None, no_probs: bool = False, no_simplify_gf: bool = False,
             unroll: int = 8, bounds: bool = False, symbolic: bool = False) -> SgclResult:
    """The host evaluator instantiated over the CPU oracle (oracle_eval.cpp): reference-order f64 arithmetic, or with
    `bounds` the reference's `--bounds` mode (TaylorPoly<Interval<F64>>, ratio constants enclosed by Number::from_ratio; the
    report prints intervals).  The interval mode is unpinned: no reference fixture runs with --bounds."""
    L = lib()
    L.orc_run_sgcl.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.orc_sgcl_report.restype = C.c_char_p
    L.orc_sgcl_report.argtypes = [C.c_void_p]
    L.orc_sgcl_moments.argtypes = [C.c_void_p, _f64p]
    L.orc_sgcl_limit.restype = C.c_uint64
    L.orc_sgcl_limit.argtypes = [C.c_void_p]
    L.orc_sgcl_probs.argtypes = [C.c_void_p, _f64p, _f64p]
    L.orc_sgcl_free.argtypes = [C.c_void_p]
    h = C.c_void_p()
    err = C.create_string_buffer(2048)
    flags = (1 if no_probs else 0) | (2 if no_simplify_gf else 0) | (4 if bounds else 0) | (8 if symbolic else 0)
    L.orc_sgcl_moment_bounds.argtypes = [C.c_void_p, _f64p]
    L.orc_sgcl_prob_bounds.argtypes = [C.c_void_p, _f64p]
    rc = L.orc_run_sgcl(source.encode(), -1 if limit is None else int(limit), flags, unroll, C.byref(h), err, 2048)
    if rc != 0:
        raise OracleError(err.value.decode())
    try:
        m = (C.c_double * 11)()
        L.orc_sgcl_moments(h, m)
        n = int(L.orc_sgcl_limit(h))
        p, q = (C.c_double * max(n, 1))(), (C.c_double * max(n, 1))()
        L.orc_sgcl_probs(h, p, q)
        mb, pb = (C.c_double * 22)(), (C.c_double * max(2 * n, 2))()
        L.orc_sgcl_moment_bounds(h, mb)
        L.orc_sgcl_prob_bounds(h, pb)
        return SgclResult(L.orc_sgcl_report(h).decode(), list(m), list(p)[:n], list(q)[:n],
                          [(mb[2 * i], mb[2 * i + 1]) for i in range(11)], [(pb[2 * i], pb[2 * i + 1]) for i in range(n)])
    finally:
        L.orc_sgcl_free(h)


class SgclBounds:
    """Enclosures [lo, hi] of the evaluator's direct outputs (oracle_eval.cpp::orc_run_sgcl_bounds)."""

    def __init__(self, rest, total, raw_moments, probs):
        self.rest = rest                  # mass of the `rest` generating function
        self.total = total                # Z before the rest mass is added and before clamping to [0, 1]
        self.raw_moments = raw_moments    # E[X], E[X^2], E[X^3], E[X^4] (normalised by Z)
        self.probs = probs                # unnormalised p(0..limit-1)


def run_sgcl_bounds(source: str, limit: int = 0, unroll: int = 8) -> SgclBounds:
    """The host evaluator instantiated over TaylorPoly<Interval<F64>> (the arithmetic of the reference's --bounds mode,
    src/interval.rs) on the DAG the f64 path evaluates: ratio constants are the enclosures Number::from_ratio builds; no
    simplification pass.  PARITY UNPINNED: no reference fixture runs with --bounds."""
    L = lib()
    L.orc_run_sgcl_bounds.argtypes = [C.c_char_p, C.c_int64, C.c_uint64, _f64p, _f64p, C.c_char_p, C.c_size_t]
    out = (C.c_double * 12)()
    probs = (C.c_double * max(2 * limit, 2))()
    err = C.create_string_buffer(2048)
    rc = L.orc_run_sgcl_bounds(source.encode(), int(limit), unroll, out, probs, err, 2048)
    if rc != 0:
        raise OracleError(err.value.decode())
    o = list(out)
    return SgclBounds((o[0], o[1]), (o[2], o[3]), [(o[4 + 2 * i], o[5 + 2 * i]) for i in range(4)],
                      [(probs[2 * i], probs[2 * i + 1]) for i in range(limit)])


def mul_macs(xshape, yshape, rshape) -> float:
    return lib().orc_mul_macs(len(rshape), _u64(xshape), _u64(yshape), _u64(rshape))
