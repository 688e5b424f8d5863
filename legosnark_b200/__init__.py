"""legosnark_b200 — B200-native MSM / fixed-base batch_exp engine behind libff's
scalar_multiplication API, as used by LegoSNARK.

This module is the Python host-side mirror of the reference interface
(LFF/algebra/scalar_multiplication/multiexp.hpp, LFF = depends/libsnark/depends/
libff/libff): same function names, argument meaning and zero/edge behaviour,
over numpy ``uint64`` arrays holding the reference's in-memory objects
(Montgomery limbs, R = 2^256):

    Fr scalars  (n, 4)      G1 points (n, 12) = X|Y|Z      G2 points (n, 24)

All arithmetic happens in ``libb200msm.so`` (CUDA, sm_100a) through the C-ABI in
``include/b200_msm.h``; the C++ drop-in header lives in ``legosnark_b200/shim``.
There is no CPU fallback: importing works anywhere, but every compute call
raises ``B200Error`` without the built library and a CUDA device.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

__all__ = [
    "B200Error", "init", "init_devices", "shutdown", "device_count", "lib", "library_path",
    "multi_exp", "multi_exp_with_mixed_addition", "multi_exp_batch", "kc_multi_exp", "get_exp_window_size", "get_window_table", "batch_exp",
    "batch_exp_with_coeff", "batch_to_special", "CommitmentKey", "sum_partials", "shard_range", "WindowTable", "last_stats", "set_tuning", "set_pipeline_chunks",
    "imad_peak", "test_field_op", "test_group_op",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libb200msm.so")
_u64p = ctypes.POINTER(ctypes.c_uint64)
_LIMBS = {"g1": 12, "g2": 24}
_AFFINE_LIMBS = {"g1": 8, "g2": 16}

# multi_exp_method (multiexp.hpp:20-47): every method computes the same group element; the engine
# runs its own Pippenger for all of them.
multi_exp_method_naive = 0
multi_exp_method_naive_plain = 1
multi_exp_method_bos_coster = 2
multi_exp_method_BDLO12 = 3


class B200Error(RuntimeError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint64), ("window_bits", ctypes.c_uint32), ("num_windows", ctypes.c_uint32),
                ("chunk_len", ctypes.c_uint32), ("kernel_launches", ctypes.c_uint32), ("num_tasks", ctypes.c_uint64),
                ("host_finalize_us", ctypes.c_double), ("h2d_bytes", ctypes.c_double), ("d2h_bytes", ctypes.c_double),
                ("num_entries", ctypes.c_uint64), ("accumulate_ms", ctypes.c_double), ("device_ms", ctypes.c_double),
                ("sort_ms", ctypes.c_double)]


_lib = None


def library_path() -> str:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the C-ABI library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise B200Error(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). legosnark_b200 has no CPU fallback.")
        L = ctypes.CDLL(_LIB_PATH)
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_version.restype = ctypes.c_char_p
        L.b200_exp_window_size_g1.restype = ctypes.c_size_t
        L.b200_exp_window_size_g1.argtypes = [ctypes.c_size_t]
        L.b200_exp_window_size_g2.restype = ctypes.c_size_t
        L.b200_exp_window_size_g2.argtypes = [ctypes.c_size_t]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise B200Error(f"{what} failed (code {rc}): {lib().b200_last_error().decode()}")


def _arr(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, width)


def _p(a):
    return None if a is None else a.ctypes.data_as(_u64p)


def _sz(n):
    return ctypes.c_size_t(int(n))


def _vp(x):
    return ctypes.c_void_p(int(x)) if x else ctypes.c_void_p(0)


# ---- lifecycle -------------------------------------------------------------------
def init(n_gpus: int = 1):
    _check(lib().b200_init(int(n_gpus)), "b200_init")


def init_async(n_gpus: int = 1):
    """b200_init on a background thread (hides the CUDA driver's start-up under the caller's own set-up work); every later
    call waits for it."""
    _check(lib().b200_init_async(int(n_gpus)), "b200_init_async")


def init_devices(ids):
    arr = (ctypes.c_int * len(ids))(*ids)
    _check(lib().b200_init_devices(arr, len(ids)), "b200_init_devices")


def shutdown():
    if _lib is not None:
        _lib.b200_shutdown()


def device_count() -> int:
    return int(lib().b200_device_count())


# ---- multi_exp / multi_exp_with_mixed_addition (multiexp.hpp:56-75) ----------------
def multi_exp(group, bases, scalars, chunks: int = 1, method: int = multi_exp_method_BDLO12):
    """sum_i scalars[i] * bases[i]; returns the normalised point (12 or 24 limbs).

    ``chunks`` and ``method`` are accepted for signature parity (multiexp.tcc:402-441);
    they never change the group element.  ``len(bases) == 0`` returns zero like the reference."""
    L = _LIMBS[group]
    bases, scalars = _arr(bases, L), _arr(scalars, 4)
    if bases.shape[0] != scalars.shape[0]:
        raise ValueError("bases and scalars differ in length")  # assert at multiexp.tcc:450
    out = np.zeros(L, dtype=np.uint64)
    _check(getattr(lib(), "b200_msm_" + group)(_p(bases), _p(scalars), _sz(bases.shape[0]), _p(out)), "b200_msm_" + group)
    return out


def multi_exp_batch(group, bases, scalars, offsets):
    """Many small MSMs in one call: out[j] = sum over [offsets[j], offsets[j+1]) of scalars[i] * bases[i].
    mtxmultiexp's per-column simplesparsemexp calls (LS/gadgets/subspace.cc:18-25, LS/utils/sparsemexp.h:62-90)
    submitted together; returns (count, limbs) normalised points."""
    L = _LIMBS[group]
    bases, scalars = _arr(bases, L), _arr(scalars, 4)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64).reshape(-1)
    if bases.shape[0] != scalars.shape[0] or offsets.shape[0] < 1 or int(offsets[-1]) != bases.shape[0]:
        raise ValueError("bases, scalars and offsets disagree")
    count = offsets.shape[0] - 1
    out = np.zeros((count, L), dtype=np.uint64)
    _check(getattr(lib(), "b200_msm_batch_" + group)(_p(bases), _p(scalars), _p(offsets), _sz(count), _p(out)), "b200_msm_batch_" + group)
    return out


def kc_multi_exp(g2_bases, g1_bases, scalars):
    """knowledge_commitment<G2,G1> MSM (SNK/knowledge_commitment/kc_multiexp.tcc:21-89): the pair
    (sum s_i g_i in G2, sum s_i h_i in G1) over one scalar vector, uploaded once."""
    g2b, g1b, scalars = _arr(g2_bases, 24), _arr(g1_bases, 12), _arr(scalars, 4)
    if not (g2b.shape[0] == g1b.shape[0] == scalars.shape[0]):
        raise ValueError("bases and scalars differ in length")
    o2, o1 = np.zeros(24, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    _check(lib().b200_msm_g2g1(_p(g2b), _p(g1b), _p(scalars), _sz(scalars.shape[0]), _p(o2), _p(o1)), "b200_msm_g2g1")
    return o2, o1


def multi_exp_with_mixed_addition(group, bases, scalars, chunks: int = 1, method: int = multi_exp_method_BDLO12):
    """multiexp.tcc:443-496 pre-filters scalars 0 and 1 on the CPU before calling multi_exp; on the
    GPU zero digits are skipped and one-digits land in bucket 1, so both entry points share a kernel."""
    return multi_exp(group, bases, scalars, chunks, method)


def sum_partials(group, pts):
    """Host-side sum of per-GPU / per-rank partial results (multiexp.tcc:433-438); normalised."""
    L = _LIMBS[group]
    pts = _arr(pts, L)
    out = np.zeros(L, dtype=np.uint64)
    _check(getattr(lib(), "b200_sum_partials_" + group)(_p(pts), _sz(pts.shape[0]), _p(out)), "b200_sum_partials")
    return out


def shard_range(n: int, rank: int, world: int):
    """Index range of `rank` when n points are split over `world` GPUs: [rank*floor(n/world), ...),
    the last rank takes the remainder (multiexp.tcc:417-431)."""
    if world <= 1 or n < world:
        return (0, n) if rank == 0 else (n, n)
    one = n // world
    lo = rank * one
    return (lo, n if rank == world - 1 else lo + one)


class CommitmentKey:
    """Device-resident bases (CommScheme's g1s / g2s, LS/prototools/commit.h:129-139)."""

    def __init__(self, group, bases=None, device_affine_ptr=None, n=None):
        self.group = group
        h = ctypes.c_uint64(0)
        if device_affine_ptr is not None:
            self.n = int(n)
            _check(getattr(lib(), "b200_pin_affine_dev_" + group)(_vp(device_affine_ptr), _sz(self.n), ctypes.byref(h)),
                   "b200_pin_affine_dev_" + group)
        else:
            bases = _arr(bases, _LIMBS[group])
            self.n = bases.shape[0]
            _check(getattr(lib(), "b200_pin_bases_" + group)(_p(bases), _sz(self.n), ctypes.byref(h)), "b200_pin_bases_" + group)
        self.handle = h.value

    def precompute(self, window_bits: int = 0):
        """Extend the key by its window multiples 2^(c k) P_i (b200_key_precompute_*): later
        multi_exps sort all windows into one bucket set.  One-off cost per key."""
        _check(getattr(lib(), "b200_key_precompute_" + self.group)(ctypes.c_uint64(self.handle), ctypes.c_uint32(int(window_bits))),
               "b200_key_precompute_" + self.group)
        return self

    def multi_exp(self, scalars, offset: int = 0):
        scalars = _arr(scalars, 4)
        out = np.zeros(_LIMBS[self.group], dtype=np.uint64)
        _check(getattr(lib(), "b200_msm_pinned_" + self.group)(ctypes.c_uint64(self.handle), _sz(offset), _p(scalars),
                                                              _sz(scalars.shape[0]), _p(out)), "b200_msm_pinned")
        return out

    def multi_exp_device(self, d_scalars_ptr: int, n: int, offset: int = 0, stream: int = 0):
        out = np.zeros(_LIMBS[self.group], dtype=np.uint64)
        _check(getattr(lib(), "b200_msm_pinned_dev_" + self.group)(ctypes.c_uint64(self.handle), _sz(offset), _vp(d_scalars_ptr),
                                                                  _sz(n), _vp(stream), _p(out)), "b200_msm_pinned_dev")
        return out

    def close(self):
        if self.handle and _lib is not None:
            _lib.b200_unpin_bases(ctypes.c_uint64(self.handle))
        self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- fixed-base exponentiation (multiexp.hpp:96-124) -----------------------------------
def get_exp_window_size(group, num_scalars: int) -> int:
    """libff's tuned table (alt_bn128_init.cpp:157-201, 220-264); kept for API parity — the engine
    picks its own window for the GPU table."""
    return int(getattr(lib(), "b200_exp_window_size_" + group)(int(num_scalars)))


class WindowTable:
    """Stands in for libff::window_table<T>: a handle to an affine table in HBM."""

    def __init__(self, group, g, expected_scalars: int = 1 << 16):
        self.group = group
        g = _arr(g, _LIMBS[group])
        h = ctypes.c_uint64(0)
        _check(getattr(lib(), "b200_window_table_create_" + group)(_p(g), _sz(expected_scalars), ctypes.byref(h)),
               "b200_window_table_create")
        self.handle = h.value

    def close(self):
        if self.handle and _lib is not None:
            _lib.b200_window_table_destroy(ctypes.c_uint64(self.handle))
        self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_window_table(group, scalar_size: int, window: int, g, expected_scalars: int = 1 << 16) -> WindowTable:
    """multiexp.tcc:547-583.  scalar_size/window are accepted for parity; no caller can observe the layout."""
    return WindowTable(group, g, expected_scalars)


def batch_exp(scalar_size: int, window: int, table: WindowTable, v):
    """multiexp.tcc:614-646: [v_i * g]; normalised points (batch_to_special form)."""
    return batch_exp_with_coeff(scalar_size, window, table, None, v)


def batch_exp_with_coeff(scalar_size: int, window: int, table: WindowTable, coeff, v):
    """multiexp.tcc:648-681: [(coeff * v_i) * g]."""
    v = _arr(v, 4)
    coeff = None if coeff is None else _arr(coeff, 4)
    out = np.zeros((v.shape[0], _LIMBS[table.group]), dtype=np.uint64)
    _check(getattr(lib(), "b200_batch_exp_table_" + table.group)(ctypes.c_uint64(table.handle), _p(v), _sz(v.shape[0]), _p(coeff),
                                                                _p(out)), "b200_batch_exp_table")
    return out


def batch_exp_device(table: WindowTable, d_scalars_ptr: int, n: int, d_out_affine_ptr: int, stream: int = 0):
    _check(getattr(lib(), "b200_batch_exp_table_dev_" + table.group)(ctypes.c_uint64(table.handle), _vp(d_scalars_ptr), _sz(n),
                                                                    _vp(d_out_affine_ptr), _vp(stream)), "b200_batch_exp_table_dev")


def batch_exp_once(group, g, v, coeff=None):
    """get_window_table + batch_exp in one call (LS/utils/util.h:119-134 simpleBatchExp)."""
    g, v = _arr(g, _LIMBS[group]), _arr(v, 4)
    coeff = None if coeff is None else _arr(coeff, 4)
    out = np.zeros((v.shape[0], _LIMBS[group]), dtype=np.uint64)
    _check(getattr(lib(), "b200_batch_exp_" + group)(_p(g), _p(v), _sz(v.shape[0]), _p(coeff), _p(out)), "b200_batch_exp")
    return out


def batch_to_special(group, vec):
    """multiexp.tcc:683-715: every point to (X/Z^2, Y/Z^3, 1); zeros to the group's zero."""
    vec = _arr(vec, _LIMBS[group]).copy()
    _check(getattr(lib(), "b200_batch_to_affine_" + group)(_p(vec), _sz(vec.shape[0])), "b200_batch_to_affine")
    return vec


# ---- Fr vector work next to the MSMs ------------------------
def fold_witness(v, r):
    """CPPoly::prove's folding (LS/gadgets/poly.h:45-67): (w_coeffs (2^d, 4), last tmp_v[0])."""
    v, r = _arr(v, 4), _arr(r, 4)
    d = r.shape[0]
    if v.shape[0] != 1 << d:
        raise ValueError("v must hold 2^d values")
    w = np.zeros((1 << d, 4), dtype=np.uint64)
    ev = np.zeros(4, dtype=np.uint64)
    _check(lib().b200_fr_fold_witness(_p(v), _p(r), _sz(d), _p(w), _p(ev)), "b200_fr_fold_witness")
    return w, ev


def evalMLE(v, r):
    """MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234)."""
    v, r = _arr(v, 4), _arr(r, 4)
    d = r.shape[0]
    if v.shape[0] != 1 << d:
        raise ValueError("assert(N == 1 << d)")  # polytools.h:211
    out = np.zeros(4, dtype=np.uint64)
    _check(lib().b200_fr_eval_mle(_p(v), _p(r), _sz(d), _p(out)), "b200_fr_eval_mle")
    return out


def mle_push_randomness(table, r):
    """DPMle::pushRandomness (LS/prototools/mle.h:199-210) on a table of 2 * half values."""
    table, r = _arr(table, 4), _arr(r, 4)
    half = table.shape[0] // 2
    out = np.zeros((half, 4), dtype=np.uint64)
    _check(lib().b200_fr_mle_bind(_p(table), _sz(half), _p(r), _p(out)), "b200_fr_mle_bind")
    return out


def fr_step_fft(a, log_big, log_small, mode=0, g=None):
    """libfqfft's step_radix2_domain (step_radix2_domain.tcc:38-152) over 2^log_big + 2^log_small points: mode 0 FFT,
    1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g); returns the transformed copy."""
    a = _arr(a, 4).copy()
    if a.shape[0] != (1 << log_big) + (1 << log_small):
        raise ValueError("step_radix2: expected a.size() == this->m")
    g = None if g is None else _arr(g, 4)
    _check(lib().b200_fr_step_fft(_p(a), _sz(log_big), _sz(log_small), int(mode), _p(g)), "b200_fr_step_fft")
    return a


def qap_h_coefficients(aA, aB, aC, log_big, log_small, g, div):
    """The vector part of libsnark's r1cs_to_qap_witness_map for d1 = d2 = d3 = 0 (r1cs_to_qap.tcc:232-311): evaluations of
    A, B, C on the domain -> coefficients of H = (A B - C) / Z.  log_small = None: basic radix-2 domain of 2^log_big points
    (div = [Z^-1]); else the step domain of 2^log_big + 2^log_small points (div = [c1, ratio, c0, Z1^-1])."""
    aA, aB, aC = _arr(aA, 4), _arr(aB, 4), _arr(aC, 4)
    m = (1 << log_big) + (0 if log_small is None else 1 << log_small)
    if not (aA.shape[0] == aB.shape[0] == aC.shape[0] == m):
        raise ValueError("expected m values per vector")
    g, div = _arr(g, 4), _arr(div, 4)
    H = np.zeros((m, 4), dtype=np.uint64)
    ls = ctypes.c_size_t(2 ** 64 - 1) if log_small is None else _sz(log_small)
    _check(lib().b200_qap_h_coefficients(_p(aA), _p(aB), _p(aC), _sz(log_big), ls, _p(g), _p(div), _p(H)), "b200_qap_h_coefficients")
    return H


def scale_inv_geometric(P, n_geo, c1, ratio, c0, tail=None):
    """P[i] *= (c1 * ratio^i - c0)^-1 for i < n_geo and P[n_geo + i] *= tail: the divisions of libfqfft's
    step_radix2_domain::divide_by_Z_on_coset (step_radix2_domain.tcc:213-241) with caller-formed constants."""
    P = _arr(P, 4).copy()
    n_tail = P.shape[0] - n_geo
    if n_tail < 0 or (n_tail and tail is None):
        raise ValueError("bad sizes")
    c1, ratio, c0 = _arr(c1, 4), _arr(ratio, 4), _arr(c0, 4)
    tail = None if tail is None else _arr(tail, 4)
    _check(lib().b200_fr_scale_inv_geometric(_p(P), _sz(n_geo), _p(c1), _p(ratio), _p(c0), _sz(n_tail), _p(tail)),
           "b200_fr_scale_inv_geometric")
    return P


def geometric_quotients(n, consts, inp=None):
    """out[i] = inp[i] * a0 * a_ratio^i / prod_f (c1_f * ratio_f^i - c0_f), i < n (b200_fr_geometric_quotients); consts = rows
    a0, a_ratio, then (c1, ratio, c0) per factor (one or two factors), Montgomery-form Fr."""
    consts = _arr(consts, 4)
    nf, rem = divmod(consts.shape[0] - 2, 3)
    if rem or nf not in (1, 2):
        raise ValueError("consts must hold 2 + 3 * n_factors rows")
    out = np.zeros((n, 4), dtype=np.uint64)
    if inp is not None:
        inp = _arr(inp, 4)
        if inp.shape[0] != n:
            raise ValueError("inp must hold n values")
    _check(lib().b200_fr_geometric_quotients(_p(out), None if inp is None else _p(inp), _sz(n), _p(consts), _sz(nf)),
           "b200_fr_geometric_quotients")
    return out


_FR_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # alt_bn128_init.cpp:40
_FR_ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904  # Fr::root_of_unity, alt_bn128_init.cpp:60


def _fr_rows(values):
    """Python integers -> Montgomery-form limb rows (R = 2^256, fp.hpp:42)."""
    out = np.zeros((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = (v % _FR_MOD) * (1 << 256) % _FR_MOD
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def _fr_int(row):
    m = sum(int(row[k]) << (64 * k) for k in range(4))
    return m * pow(1 << 256, -1, _FR_MOD) % _FR_MOD


def evaluate_all_lagrange_polynomials(log_big, log_small, t):
    """evaluate_all_lagrange_polynomials(t) of libfqfft's basic_radix2_domain (log_small None; basic_radix2_domain_aux.tcc:183-236)
    or step_radix2_domain (2^log_big + 2^log_small points; step_radix2_domain.tcc:161-186).  The host forms the constants the way
    the shadow headers do (shim/libfqfft/.../basic_radix2_domain_aux.hpp, step_radix2_domain.hpp); the per-point divisions run on
    the device.  t on the domain itself gives the reference's unit vector (aux.tcc:201-214)."""
    r = _FR_MOD
    tv = _fr_int(_arr(t, 4).reshape(-1, 4)[0])

    def basic(log_m, tt, scale, extra=None):
        m = 1 << log_m
        if m == 1:
            return _fr_rows([scale])
        omega = pow(_FR_ROOT_2_28, 1 << (28 - log_m), r)
        tm = pow(tt, m, r)
        if tm == 1:  # t is omega^k: 1 at k, 0 elsewhere (times what the step domain multiplies in)
            out = np.zeros((m, 4), dtype=np.uint64)
            k = next(i for i in range(m) if pow(omega, i, r) == tt)
            f = scale
            if extra is not None:
                c1, ratio, c0 = extra
                f = f * pow((c1 * pow(ratio, k, r) - c0) % r, -1, r) % r
            out[k] = _fr_rows([f])[0]
            return out
        l0 = (tm - 1) * pow(m, -1, r) % r * scale % r
        rows = [l0, omega, r - 1, omega, (r - tt) % r]  # l0 omega^i / (t - omega^i)
        if extra is not None:
            rows += list(extra)
        return geometric_quotients(m, _fr_rows(rows))

    if log_small is None:
        return basic(log_big, tv, 1)
    big, small = 1 << log_big, 1 << log_small
    omega = pow(_FR_ROOT_2_28, 1 << (28 - (log_big + 1)), r)
    big_omega = omega * omega % r
    w = pow(omega, small, r)
    L0 = (pow(tv, small, r) - w) % r
    L1 = (pow(tv, big, r) - 1) * pow((pow(omega, big, r) - 1) % r, -1, r) % r
    head = basic(log_big, tv, L0, (1, pow(big_omega, small, r), w))  # inner_big[i] * L0 / (rho^i - omega^small)
    tail = basic(log_small, tv * pow(omega, -1, r) % r, L1)
    return np.concatenate([head, tail])


def compute_eq_tbl(r):
    """DPBeta::compute_eq_tbl (LS/prototools/mle.h:93-105): the 2^d-entry table, level by level as the reference writes it."""
    r = _arr(r, 4)
    d = r.shape[0]
    out = np.zeros((1 << d, 4), dtype=np.uint64)
    _check(lib().b200_fr_eq_table(_p(r), _sz(d), _p(out)), "b200_fr_eq_table")
    return out


def matrix_mle(A, rho):
    """DPMatrixMle's constructor (LS/prototools/mle.h:241-259): v[r] = sum_l A[(l << d) + r] * eq(rho)[l]."""
    A, rho = _arr(A, 4), _arr(rho, 4)
    d = rho.shape[0]
    if A.shape[0] != 1 << (2 * d):
        raise ValueError("A must hold 2^d x 2^d values")
    out = np.zeros((1 << d, 4), dtype=np.uint64)
    _check(lib().b200_fr_matrix_mle(_p(A), _p(rho), _sz(d), _p(out)), "b200_fr_matrix_mle")
    return out


def sumcheck_round(a, b, w=None):
    """The sum over p inside CPSumcheck::make_new_h_poly (LS/gadgets/sumcheck.h:85-106) for two DPMle tables of
    2 * half values; w = beta suffix values or None (DPBetaDummy).  Returns the (3, 4) coefficients."""
    a, b = _arr(a, 4), _arr(b, 4)
    half = a.shape[0] // 2
    if b.shape[0] != 2 * half or (w is not None and _arr(w, 4).shape[0] != half):
        raise ValueError("table sizes disagree")
    w = None if w is None else _arr(w, 4)
    out = np.zeros((3, 4), dtype=np.uint64)
    _check(lib().b200_fr_sumcheck_round(_p(a), _p(b), _p(w), _sz(half), _p(out)), "b200_fr_sumcheck_round")
    return out


def sumcheck_rounds(a, b, r):
    """All d round polynomials of the beta-less sum-check (CPSumcheckMatrix; the loop of CPSumcheck::prove,
    LS/gadgets/sumcheck.cc:56-70), tables bound to r[i] on the device between rounds.  Returns (d, 3, 4)."""
    a, b, r = _arr(a, 4), _arr(b, 4), _arr(r, 4)
    d = r.shape[0]
    if a.shape[0] != 1 << d or b.shape[0] != 1 << d:
        raise ValueError("tables must hold 2^d values")
    out = np.zeros((d, 3, 4), dtype=np.uint64)
    _check(lib().b200_fr_sumcheck_rounds(_p(a), _p(b), _p(r), _sz(d), _p(out)), "b200_fr_sumcheck_rounds")
    return out


def cppoly_prove(key: "CommitmentKey", v, r):
    """CPPoly::prove (LS/gadgets/poly.h:45-91) against a resident G1 key: (witness (d, 12), evalMLE(v, r))."""
    v, r = _arr(v, 4), _arr(r, 4)
    d = r.shape[0]
    if v.shape[0] != 1 << d:
        raise ValueError("v must hold 2^d values")
    wit = np.zeros((d, 12), dtype=np.uint64)
    ev = np.zeros(4, dtype=np.uint64)
    _check(lib().b200_cppoly_prove_g1(ctypes.c_uint64(key.handle), _p(v), _p(r), _sz(d), _p(wit), _p(ev)), "b200_cppoly_prove_g1")
    return wit, ev


FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3


def fr_fft(a, mode: int = FFT, g=None):
    """libfqfft basic_radix2_domain<Fr>::FFT / iFFT / cosetFFT / icosetFFT on 2^k values; returns the transform."""
    a = _arr(a, 4).copy()
    n = a.shape[0]
    log_n = n.bit_length() - 1
    if n != 1 << log_n:
        raise ValueError("basic_radix2: expected a power-of-two size")
    g = None if g is None else _arr(g, 4)
    _check(lib().b200_fr_fft(_p(a), _sz(log_n), int(mode), None if g is None else _p(g)), "b200_fr_fft")
    return a


def fr_fft_device(d_ptr: int, log_n: int, mode: int = FFT, g=None, stream: int = 0):
    g = None if g is None else _arr(g, 4)
    _check(lib().b200_fr_fft_dev(_vp(d_ptr), _sz(log_n), int(mode), None if g is None else _p(g), _vp(stream)), "b200_fr_fft_dev")


# ---- wire format: point compression --------------------------------
ALT_BN128, ALT_BN128_MONTGOMERY_OUTPUT, BN128 = 0, 1, 2
_u8p = ctypes.POINTER(ctypes.c_uint8)


def compress_points(group, pts, flavour: int = ALT_BN128):
    """The arithmetic of the reference's compressed operator<< (alt_bn128_g1.cpp:404-419, bn128_g1.cpp:344-373):
    (X in wire form (n, 4|8), flags (n,) uint8: bit 0 = Y bit, bit 1 = is_zero)."""
    pts = _arr(pts, _LIMBS[group])
    n = pts.shape[0]
    x = np.zeros((n, _LIMBS[group] // 3), dtype=np.uint64)
    flags = np.zeros(n, dtype=np.uint8)
    _check(getattr(lib(), "b200_compress_" + group)(_p(pts), _sz(n), int(flavour), _p(x), flags.ctypes.data_as(_u8p)),
           "b200_compress_" + group)
    return x, flags


def decompress_points(group, x, flags, flavour: int = ALT_BN128, report_bad: bool = False):
    """operator>> with point compression (alt_bn128_g1.cpp:421-459): (X, Y, 1) with Y = +-sqrt(X^3 + b), or the zero."""
    x = _arr(x, _LIMBS[group] // 3)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    n = x.shape[0]
    out = np.zeros((n, _LIMBS[group]), dtype=np.uint64)
    bad = np.zeros(n, dtype=np.uint8) if report_bad else None
    _check(getattr(lib(), "b200_decompress_" + group)(_p(x), flags.ctypes.data_as(_u8p), _sz(n), int(flavour), _p(out),
                                                      None if bad is None else bad.ctypes.data_as(_u8p)), "b200_decompress_" + group)
    return (out, bad) if report_bad else out


# ---- introspection ---------------------------------------------------------------------
def last_stats() -> dict:
    s = Stats()
    lib().b200_last_stats(ctypes.byref(s))
    return {k: getattr(s, k) for k, _ in Stats._fields_}


def set_tuning(window_bits: int = 0, chunk_len: int = 0):
    _check(lib().b200_set_tuning(int(window_bits), int(chunk_len)), "b200_set_tuning")


def set_tuning_ex(key: str, value: int):
    """Sweep knobs: reduce_log_segment (-1 = model), reduce_split (0 = auto), use_precomputed (0/1)."""
    _check(lib().b200_set_tuning_ex(key.encode(), int(value)), "b200_set_tuning_ex")


def set_pipeline_chunks(chunks: int = 0):
    """Upload chunks of a host-buffer multi_exp (0 = auto, 1 = no pipelining)."""
    _check(lib().b200_set_pipeline_chunks(int(chunks)), "b200_set_pipeline_chunks")


def imad_peak(kind: int = 0, iters: int = 4096):
    ops = ctypes.c_double(0)
    ms = ctypes.c_double(0)
    _check(lib().b200_imad_peak(int(kind), int(iters), ctypes.byref(ops), ctypes.byref(ms)), "b200_imad_peak")
    return ops.value, ms.value


def test_field_op(field: int, op: int, a, b=None):
    W = 8 if field == 2 else 4
    a = _arr(a, W)
    b = None if b is None else _arr(b, W)
    out = np.zeros_like(a)
    _check(lib().b200_test_field_op(int(field), int(op), _p(a), _p(b), _sz(a.shape[0]), _p(out)), "b200_test_field_op")
    return out


def test_group_op(group: int, op: int, a, b=None, k: int = 0):
    L = 12 if group == 0 else 24
    a = _arr(a, L)
    b = None if b is None else _arr(b, L)
    out = np.zeros_like(a)
    _check(lib().b200_test_group_op(int(group), int(op), _p(a), _p(b), _sz(a.shape[0]), ctypes.c_uint32(k), _p(out)),
           "b200_test_group_op")
    return out
