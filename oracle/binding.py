"""ctypes bindings for the two parity checkers.  TEST INFRASTRUCTURE ONLY.

* ``Checker("orc")`` -> oracle/libbn254oracle.so, the plain-C restatement
  (oracle/bn254_oracle.c).
* ``Checker("ref")`` -> oracle/_ref/libffref.so, the UNMODIFIED reference libff
  sources compiled in place by oracle/Makefile (wrapper: oracle/ref_wrap.cpp).

Both export the same functions (prefix ``orc_`` / ``ref_``); arrays are numpy
``uint64`` little-endian Montgomery limbs: Fr/Fq ``(n, 4)``, Fq2 ``(n, 8)``,
G1 ``(n, 12)`` Jacobian X|Y|Z, G2 ``(n, 24)``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_PATH = os.path.join(HERE, "libbn254oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libffref.so")
REF_NATIVE_PATH = os.path.join(HERE, "_ref", "libffref_native.so")  # same sources, -march=native (timed baseline only)
LSREF_PATH = os.path.join(HERE, "_ref", "liblsref.so")  # oracle/ref_wrap_ls.cpp: LegoSNARK / libfqfft Fr-side routines

_u64p = ctypes.POINTER(ctypes.c_uint64)

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MONT_R = 1 << 256


def build_oracle() -> str:
    """Compile the C restatement (gcc only; works on the GPU box too)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return ORC_PATH


def build_ref() -> str | None:
    """Compile oracle/_ref from /root/reference when it is present."""
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
    return REF_PATH if os.path.exists(REF_PATH) else None


def native_build_runs() -> bool:
    """The -march=native build was compiled in the build container; run a tiny MSM with it in a child process
    so that an instruction this host lacks ends the child, not the caller."""
    code = ("import ctypes, numpy as np; L = ctypes.CDLL(%r); L.ref_init(); o = np.zeros(12, dtype=np.uint64); "
            "b = np.zeros((4, 12), dtype=np.uint64); s = np.zeros((4, 4), dtype=np.uint64); "
            "p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)); "
            "raise SystemExit(L.ref_msm_g1(0, p(b), p(s), ctypes.c_size_t(4), ctypes.c_size_t(1), 0, 1, p(o)))" % REF_NATIVE_PATH)
    try:
        import sys
        return subprocess.run([sys.executable, "-c", code], capture_output=True, timeout=120).returncode == 0
    except Exception:
        return False


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_u64p)


def _c(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if width is not None:
        a = a.reshape(-1, width)
    return a


class Checker:
    """Uniform view over the restatement ("orc") and the compiled reference ("ref")."""

    def __init__(self, kind: str = "orc", native: bool = False):
        assert kind in ("orc", "ref")
        self.kind = kind
        self.native = False
        if kind == "orc":
            if not os.path.exists(ORC_PATH) or os.path.getmtime(ORC_PATH) < os.path.getmtime(
                os.path.join(HERE, "bn254_oracle.c")
            ):
                build_oracle()
            path = ORC_PATH
        else:
            if not os.path.exists(REF_PATH):
                if os.path.isdir("/root/reference"):
                    build_ref()
            if not os.path.exists(REF_PATH):
                raise FileNotFoundError("oracle/_ref/libffref.so not built (needs /root/reference)")
            path = REF_PATH
            if native and os.path.exists(REF_NATIVE_PATH) and native_build_runs():
                path, self.native = REF_NATIVE_PATH, True
        self.lib = ctypes.CDLL(path)
        self.p = kind + "_"
        self.curve_arg = kind == "ref"
        if kind == "ref":
            self.lib.ref_init()
        for name in ("exp_window_size_g1", "exp_window_size_g2"):
            f = getattr(self.lib, self.p + name)
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_size_t]

    @staticmethod
    def available(kind: str) -> bool:
        if kind == "orc":
            return True
        return os.path.exists(REF_PATH) or os.path.isdir("/root/reference")

    # -- helpers ---------------------------------------------------------
    def _call(self, name, *args):
        rc = getattr(self.lib, self.p + name)(*args)
        if rc != 0:
            raise RuntimeError(f"{self.p}{name} returned {rc}")

    def max_threads(self) -> int:
        return int(getattr(self.lib, self.p + "max_threads")())

    # -- MSM ---------------------------------------------------------------
    def msm(self, group, bases, scalars, chunks=1, variant=0, normalise=True, curve=0):
        L = 12 if group == "g1" else 24
        bases = _c(bases, L)
        scalars = _c(scalars, 4)
        n = bases.shape[0]
        assert scalars.shape[0] == n
        out = np.zeros(L, dtype=np.uint64)
        args = [_ptr(bases), _ptr(scalars), ctypes.c_size_t(n), ctypes.c_size_t(chunks), int(variant),
                int(bool(normalise)), _ptr(out)]
        if self.curve_arg:
            args = [int(curve)] + args
        self._call("msm_" + group, *args)
        return out

    # -- fixed-base batch_exp ---------------------------------------------
    def batch_exp(self, group, base, scalars, coeff=None, window=0, normalise=True, curve=0):
        L = 12 if group == "g1" else 24
        base = _c(base, L)
        scalars = _c(scalars, 4)
        n = scalars.shape[0]
        coeff = None if coeff is None else _c(coeff, 4)
        out = np.zeros((n, L), dtype=np.uint64)
        args = [_ptr(base), _ptr(scalars), ctypes.c_size_t(n), _ptr(coeff), ctypes.c_size_t(window),
                int(bool(normalise)), _ptr(out)]
        if self.curve_arg:
            args = [int(curve)] + args
        self._call("batch_exp_" + group, *args)
        return out

    def exp_window_size(self, group, n) -> int:
        return int(getattr(self.lib, self.p + "exp_window_size_" + group)(n))

    def batch_to_special(self, group, pts, curve=0):
        L = 12 if group == "g1" else 24
        pts = _c(pts, L).copy()
        args = [_ptr(pts), ctypes.c_size_t(pts.shape[0])]
        if self.curve_arg:
            args = [int(curve)] + args
        self._call("batch_to_special_" + group, *args)
        return pts

    # -- element-wise group / field ops ------------------------------------
    def group_op(self, group, op, a, b=None, curve=0):
        L = 12 if group == "g1" else 24
        a = _c(a, L)
        b = None if b is None else _c(b, L)
        out = np.zeros_like(a)
        args = [int(op), _ptr(a), _ptr(b), ctypes.c_size_t(a.shape[0]), _ptr(out)]
        if self.curve_arg:
            args = [int(curve)] + args
        self._call(group + "_op", *args)
        return out

    def scalar_mul(self, group, base, scalars, stride_base=False, normalise=True):
        L = 12 if group == "g1" else 24
        base = _c(base, L)
        scalars = _c(scalars, 4)
        n = scalars.shape[0]
        out = np.zeros((n, L), dtype=np.uint64)
        self._call("scalar_mul_" + group, _ptr(base), _ptr(scalars), ctypes.c_size_t(n), int(bool(stride_base)),
                   int(bool(normalise)), _ptr(out))
        return out

    def field_op(self, field, op, a, b=None):
        W = {"fq": 4, "fr": 4, "fq2": 8}[field]
        a = _c(a, W)
        b = None if b is None else _c(b, W)
        out = np.zeros_like(a)
        self._call(field + "_op", int(op), _ptr(a), _ptr(b), ctypes.c_size_t(a.shape[0]), _ptr(out))
        return out

    def fr_from_bigint(self, a):
        a = _c(a, 4)
        out = np.zeros_like(a)
        self._call("fr_from_bigint", _ptr(a), ctypes.c_size_t(a.shape[0]), _ptr(out))
        return out

    def fr_as_bigint(self, a):
        a = _c(a, 4)
        out = np.zeros_like(a)
        self._call("fr_as_bigint", _ptr(a), ctypes.c_size_t(a.shape[0]), _ptr(out))
        return out

    def fq_from_bigint(self, a):
        a = _c(a, 4)
        out = np.zeros_like(a)
        self._call("fq_from_bigint", _ptr(a), ctypes.c_size_t(a.shape[0]), _ptr(out))
        return out

    def sha512_rng_fr(self, idx0, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        self._call("sha512_rng_fr", ctypes.c_uint64(idx0), ctypes.c_size_t(n), _ptr(out))
        return out

    # -- Fr vector work next to the MSMs ----------------------
    def _ls(self):
        """Library holding the fr_* entry points: the restatement itself, or oracle/_ref/liblsref.so
        (the reference's poly.h / polytools.h / mle.h / libfqfft behind oracle/ref_wrap_ls.cpp)."""
        if self.kind == "orc":
            return self.lib
        if getattr(self, "_lslib", None) is None:
            if not os.path.exists(LSREF_PATH) and os.path.isdir("/root/reference"):
                build_ref()
            if not os.path.exists(LSREF_PATH):
                raise FileNotFoundError("oracle/_ref/liblsref.so not built (needs /root/reference)")
            self._lslib = ctypes.CDLL(LSREF_PATH)
            self._lslib.ref_ls_init()
        return self._lslib

    def _lscall(self, name, *args):
        rc = getattr(self._ls(), self.p + name)(*args)
        if rc != 0:
            raise RuntimeError(f"{self.p}{name} returned {rc}")

    def fr_eval_mle(self, v, r):
        v, r = _c(v, 4), _c(r, 4)
        d = r.shape[0]
        assert v.shape[0] == 1 << d
        out = np.zeros(4, dtype=np.uint64)
        self._lscall("fr_eval_mle", _ptr(v), _ptr(r), ctypes.c_size_t(d), _ptr(out))
        return out

    def fr_fold_witness(self, v, r):
        """(w_coeffs (2^d, 4), eval (4,)) of CPPoly::prove's folding; restatement only (the reference keeps
        w_coeffs local to prove(): it is pinned through cppoly_prove_g1 and eval_mle)."""
        assert self.kind == "orc"
        v, r = _c(v, 4), _c(r, 4)
        d = r.shape[0]
        assert v.shape[0] == 1 << d
        w = np.zeros((1 << d, 4), dtype=np.uint64)
        ev = np.zeros(4, dtype=np.uint64)
        self._lscall("fr_fold_witness", _ptr(v), _ptr(r), ctypes.c_size_t(d), _ptr(w), _ptr(ev))
        return w, ev

    def fr_mle_bind(self, table, r):
        table, r = _c(table, 4), _c(r, 4)
        half = table.shape[0] // 2
        out = np.zeros((half, 4), dtype=np.uint64)
        self._lscall("fr_mle_bind", _ptr(table), ctypes.c_size_t(half), _ptr(r), _ptr(out))
        return out

    def fr_step_fft(self, a, log_big, log_small, mode=0, g=None):
        """libfqfft step_radix2_domain over 2^log_big + 2^log_small points; modes as fr_fft (0..3)"""
        a = _c(a, 4).copy()
        assert a.shape[0] == (1 << log_big) + (1 << log_small)
        g = None if g is None else _c(g, 4)
        self._lscall("fr_step_fft", _ptr(a), ctypes.c_size_t(log_big), ctypes.c_size_t(log_small), int(mode), _ptr(g))
        return a

    def fr_step_divide_z(self, a, log_big, log_small):
        """reference only: step_radix2_domain::divide_by_Z_on_coset"""
        assert self.kind == "ref"
        a = _c(a, 4).copy()
        self._lscall("fr_step_divide_z", _ptr(a), ctypes.c_size_t(log_big), ctypes.c_size_t(log_small))
        return a

    def fr_lagrange(self, log_big, log_small, t):
        """evaluate_all_lagrange_polynomials(t) of libfqfft's basic_radix2_domain (log_small None: 2^log_big points) or
        step_radix2_domain (2^log_big + 2^log_small points)"""
        t = _c(t, 4)
        m = (1 << log_big) + (0 if log_small is None else 1 << log_small)
        out = np.zeros((m, 4), dtype=np.uint64)
        ls = ctypes.c_size_t(-1 if log_small is None else log_small)
        self._lscall("fr_lagrange", _ptr(out), ctypes.c_size_t(log_big), ls, _ptr(t))
        return out

    # -- sum-check tables (LS/prototools/mle.h, LS/gadgets/sumcheck.h) --------------------
    def fr_eq_table(self, r):
        r = _c(r, 4)
        d = r.shape[0]
        out = np.zeros((1 << d, 4), dtype=np.uint64)
        self._lscall("fr_eq_table", _ptr(r), ctypes.c_size_t(d), _ptr(out))
        return out

    def fr_matrix_mle(self, A, rho):
        A, rho = _c(A, 4), _c(rho, 4)
        d = rho.shape[0]
        assert A.shape[0] == 1 << (2 * d)
        out = np.zeros((1 << d, 4), dtype=np.uint64)
        self._lscall("fr_matrix_mle", _ptr(A), _ptr(rho), ctypes.c_size_t(d), _ptr(out))
        return out

    def fr_sumcheck_round(self, a, b, w=None):
        """restatement only: (3, 4) coefficients of sum_p w[p] * mle_a_poly(p) * mle_b_poly(p)"""
        assert self.kind == "orc"
        a, b = _c(a, 4), _c(b, 4)
        half = a.shape[0] // 2
        w = None if w is None else _c(w, 4)
        out = np.zeros((3, 4), dtype=np.uint64)
        self._lscall("fr_sumcheck_round", _ptr(a), _ptr(b), _ptr(w), ctypes.c_size_t(half), _ptr(out))
        return out

    def fr_sumcheck_rounds(self, a, b, r):
        """(d, 3, 4): the beta-less round polynomials.  ref: CPSumcheck::make_new_h_poly with DPBetaDummy."""
        a, b, r = _c(a, 4), _c(b, 4), _c(r, 4)
        d = r.shape[0]
        assert a.shape[0] == 1 << d
        if self.kind == "orc":
            out = np.zeros((d, 3, 4), dtype=np.uint64)
            self._lscall("fr_sumcheck_rounds", _ptr(a), _ptr(b), _ptr(r), ctypes.c_size_t(d), _ptr(out))
            return out
        h, nc = self.sumcheck_h_polys(a, b, None, r)
        assert nc == 3
        return h[:, :3]

    def sumcheck_h_polys(self, a, b, rho, r):
        """reference only: h[i] of every round through CPSumcheck::make_new_h_poly, with DPBeta(rho) or, for
        rho = None, DPBetaDummy; returns ((d, 4, 4) coefficients, coefficients per polynomial)."""
        assert self.kind == "ref"
        a, b, r = _c(a, 4), _c(b, 4), _c(r, 4)
        d = r.shape[0]
        rho = None if rho is None else _c(rho, 4)
        out = np.zeros((d, 4, 4), dtype=np.uint64)
        nc = ctypes.c_size_t(0)
        rc = self._ls().ref_sumcheck_h_polys(_ptr(a), _ptr(b), _ptr(rho), _ptr(r), ctypes.c_size_t(d), ctypes.c_size_t(4), _ptr(out),
                                             ctypes.byref(nc))
        if rc != 0:
            raise RuntimeError("ref_sumcheck_h_polys failed")
        return out, int(nc.value)

    def fr_beta_suffix(self, rho):
        """reference only: DPBeta(d, rho).beta_suff_rho_cur[0 .. 2^(d-1)) after construction (the round-0 suffix values)"""
        assert self.kind == "ref"
        rho = _c(rho, 4)
        d = rho.shape[0]
        out = np.zeros((1 << (d - 1), 4), dtype=np.uint64)
        if self._ls().ref_fr_beta_suffix(_ptr(rho), ctypes.c_size_t(d), _ptr(out)) != 0:
            raise RuntimeError("ref_fr_beta_suffix failed")
        return out

    def fr_fft(self, a, mode=0, g=None):
        a = _c(a, 4).copy()
        log_n = int(a.shape[0]).bit_length() - 1
        assert a.shape[0] == 1 << log_n
        g = None if g is None else _c(g, 4)
        self._lscall("fr_fft", _ptr(a), ctypes.c_size_t(log_n), int(mode), _ptr(g))
        return a

    def cppoly_prove_g1(self, bases, v, r):
        """CPPoly::prove's witness points (d, 12), affine-normalised.  ref: the reference class itself over
        an installed key; orc: the restated folding + the restated multi_exp_with_mixed_addition."""
        bases, v, r = _c(bases, 12), _c(v, 4), _c(r, 4)
        d = r.shape[0]
        out = np.zeros((d, 12), dtype=np.uint64)
        if self.kind == "ref":
            outa = np.zeros((d, 12), dtype=np.uint64)
            self._lscall("cppoly_prove_g1", _ptr(bases), ctypes.c_size_t(bases.shape[0]), _ptr(v), _ptr(r), ctypes.c_size_t(d),
                         _ptr(out), _ptr(outa))
            assert (out == outa).all()  # witnessa repeats the MSM over the same bases (poly.h:84-86)
            return out
        w, _ = self.fr_fold_witness(v, r)
        N = 1 << d
        for i in range(d):
            m = 1 << (d - i - 1)
            start = N - 2 * m
            out[i] = self.msm("g1", bases[:m], w[start:start + m], chunks=1, variant=1)
        return out

    # -- wire format: point compression ---------------------------------
    def compress(self, group, pts, flavour=0):
        """(x (n, 4|8), flags (n,) uint8).  ref: the reference's operator<< (flavour 0 -> alt_bn128, 2 -> bn128)."""
        L = 12 if group == "g1" else 24
        pts = _c(pts, L)
        n = pts.shape[0]
        x = np.zeros((n, L // 3), dtype=np.uint64)
        flags = np.zeros(n, dtype=np.uint8)
        fp = flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        if self.kind == "ref":
            assert flavour in (0, 2), "the reference build has alt_bn128 (flavour 0) and bn128 (flavour 2)"
            self._call("compress_" + group, 0 if flavour == 0 else 1, _ptr(pts), ctypes.c_size_t(n), _ptr(x), fp)
        else:
            self._call("compress_" + group, _ptr(pts), ctypes.c_size_t(n), int(flavour), _ptr(x), fp)
        return x, flags

    def decompress(self, group, x, flags, flavour=0):
        L = 12 if group == "g1" else 24
        x = _c(x, L // 3)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        n = x.shape[0]
        out = np.zeros((n, L), dtype=np.uint64)
        fp = flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        if self.kind == "ref":
            assert flavour in (0, 2)
            self._call("decompress_" + group, 0 if flavour == 0 else 1, _ptr(x), fp, ctypes.c_size_t(n), _ptr(out))
        else:
            self._call("decompress_" + group, _ptr(x), fp, ctypes.c_size_t(n), int(flavour), _ptr(out))
        return out

    def one(self, group):
        L = 12 if group == "g1" else 24
        out = np.zeros(L, dtype=np.uint64)
        self._call(group + "_one", _ptr(out))
        return out


# -- pure-python integer helpers shared by tests / bench (no group arithmetic) --
def int_to_limbs(x: int, n: int = 4) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def limbs_to_int(a) -> int:
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(v) << (64 * i) for i, v in enumerate(a))


def ints_to_mont(xs, mod: int) -> np.ndarray:
    """Standard integers -> (n, 4) Montgomery limbs for the field of the given modulus."""
    out = np.zeros((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i] = int_to_limbs((x % mod) * MONT_R % mod)
    return out


def mont_to_ints(a, mod: int):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(MONT_R, -1, mod)
    return [limbs_to_int(row) * rinv % mod for row in a]
