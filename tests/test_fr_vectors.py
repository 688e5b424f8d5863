"""Fr vector work next to the MSMs: CPPoly::prove's folding
(LS/gadgets/poly.h:45-67), MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234),
DPMle::pushRandomness (LS/prototools/mle.h:199-210) and libfqfft's basic radix-2 domain.

CPU part: the plain-C restatement against the fixtures the reference itself produced
(tools/make_golden_fr.py) and, in the build container, against the reference live.
GPU part (-m gpu): the CUDA kernels through the C-ABI against the same fixtures and the oracle,
plus size-independent properties at benchmark sizes (iFFT o FFT = id, linearity, fold == evalMLE)."""
import numpy as np
import pytest

from oracle.binding import R_ORDER, ints_to_mont

MODES = (0, 1, 2, 3)   # FFT, iFFT, cosetFFT, icosetFFT: what the golden file holds
ALL_MODES = MODES + (4,)  # + the unscaled inverse _basic_radix2_FFT(a, omega^-1)


def _fft_key(d, mode, second_g):
    return f"fft_{d}_m{mode}" + ("_g2" if second_g else "")


# ---------------------------------------------------------------- CPU: oracle pinned
def test_oracle_fr_vectors_golden(orc, golden):
    g = golden("fr_vectors")
    for d in g["dims"]:
        d = int(d)
        v, r = g[f"v_{d}"], g[f"r_{d}"]
        assert (orc.fr_eval_mle(v, r) == g[f"eval_mle_{d}"]).all(), d
        w, ev = orc.fr_fold_witness(v, r)
        assert (ev == g[f"eval_mle_{d}"]).all(), d          # last tmp_v[0] of prove() is evalMLE(v, r)
        assert (w[-1] == 0).all()                            # w_coeffs(1 << d): last entry never written
        assert (orc.fr_mle_bind(v, r[:1]) == g[f"mle_bind_{d}"]).all(), d
        for mode in MODES:
            gg = g["coset_g"] if mode >= 2 else None
            assert (orc.fr_fft(v, mode, gg) == g[_fft_key(d, mode, False)]).all(), (d, mode)
            if mode >= 2:
                assert (orc.fr_fft(v, mode, g["coset_g2"]) == g[_fft_key(d, mode, True)]).all(), (d, mode)


def test_oracle_cppoly_prove_golden(orc, golden):
    g = golden("fr_vectors")
    for d in g["prove_dims"]:
        d = int(d)
        got = orc.cppoly_prove_g1(g[f"prove_bases_{d}"], g[f"prove_v_{d}"], g[f"prove_r_{d}"])
        assert (got == g[f"prove_witness_{d}"]).all(), d


def test_oracle_vs_reference_live(orc, ref):
    for d in (1, 4, 9, 12):
        v = orc.sha512_rng_fr(9000 + d, 1 << d)
        r = orc.sha512_rng_fr(9100 + d, d)
        assert (orc.fr_eval_mle(v, r) == ref.fr_eval_mle(v, r)).all()
        assert (orc.fr_mle_bind(v, r[-1:]) == ref.fr_mle_bind(v, r[-1:])).all()
        g = orc.sha512_rng_fr(9200 + d, 1)
        for mode in ALL_MODES:
            assert (orc.fr_fft(v, mode, g) == ref.fr_fft(v, mode, g)).all(), (d, mode)


def test_fft_properties_oracle(orc):
    """iFFT(FFT(a)) == a, icosetFFT(cosetFFT(a)) == a, FFT of a delta is all-ones."""
    d = 8
    a = orc.sha512_rng_fr(31, 1 << d)
    g = ints_to_mont([5], R_ORDER)
    assert (orc.fr_fft(orc.fr_fft(a, 0), 1) == a).all()
    assert (orc.fr_fft(orc.fr_fft(a, 2, g), 3, g) == a).all()
    delta = np.zeros((1 << d, 4), dtype=np.uint64)
    delta[0] = ints_to_mont([1], R_ORDER)[0]
    assert (orc.fr_fft(delta, 0) == ints_to_mont([1], R_ORDER)[0]).all()


# ---------------------------------------------------------------- GPU: parity through the C-ABI
@pytest.mark.gpu
def test_gpu_fr_vectors_golden(engine, golden):
    g = golden("fr_vectors")
    for d in g["dims"]:
        d = int(d)
        v, r = g[f"v_{d}"], g[f"r_{d}"]
        assert (engine.evalMLE(v, r) == g[f"eval_mle_{d}"]).all(), d
        assert (engine.mle_push_randomness(v, r[:1]) == g[f"mle_bind_{d}"]).all(), d
        for mode in MODES:
            gg = g["coset_g"] if mode >= 2 else None
            assert (engine.fr_fft(v, mode, gg) == g[_fft_key(d, mode, False)]).all(), (d, mode)
            if mode >= 2:  # a second shift: the cached coset tables must be rebuilt
                assert (engine.fr_fft(v, mode, g["coset_g2"]) == g[_fft_key(d, mode, True)]).all(), (d, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [0, 1, 2, 5, 9, 10, 13, 16])
def test_gpu_fold_vs_oracle(engine, orc, d):
    v = orc.sha512_rng_fr(700 + d, 1 << d)
    r = orc.sha512_rng_fr(800 + d, d).reshape(-1, 4)
    w, ev = engine.fold_witness(v, r)
    ow, oev = orc.fr_fold_witness(v, r)
    assert (w == ow).all() and (ev == oev).all()
    assert (engine.evalMLE(v, r) == orc.fr_eval_mle(v, r)).all()
    if d >= 1:
        assert (engine.mle_push_randomness(v, r[:1]) == orc.fr_mle_bind(v, r[:1])).all()


@pytest.mark.gpu
@pytest.mark.parametrize("d", [1, 3, 9, 10, 11, 12, 14, 17])
def test_gpu_fft_vs_oracle(engine, orc, d):
    """Sizes on both sides of the one-pass limit (2^10) and with 2 and 3 passes."""
    a = orc.sha512_rng_fr(1700 + d, 1 << d)
    g = orc.sha512_rng_fr(1800 + d, 1)
    for mode in ALL_MODES:
        assert (engine.fr_fft(a, mode, g) == orc.fr_fft(a, mode, g)).all(), (d, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [20, 22])
def test_gpu_fft_properties_at_scale(engine, orc, d):
    """BASELINE sizes (2^20 constraints and the 128x128 matrix product's 2^21-2^22 domain): round trips,
    linearity FFT(a + b) = FFT(a) + FFT(b), and a spot check of single outputs against the DFT sum."""
    n = 1 << d
    rng = np.random.default_rng(d)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    b = np.roll(a, 1, axis=0)
    g = ints_to_mont([5], R_ORDER)
    fa = engine.fr_fft(a, 0)
    assert (engine.fr_fft(fa, 1) == a).all()
    assert (engine.fr_fft(engine.fr_fft(a, 2, g), 3, g) == a).all()
    fb = engine.fr_fft(b, 0)
    ab = engine.test_field_op(1, 2, a, b)
    assert (engine.fr_fft(ab, 0) == engine.test_field_op(1, 2, fa, fb)).all()
    # FFT(a)[0] = sum a_i, and FFT of the shifted vector: FFT(b)[k] = omega^k FFT(a)[k] at k = n/2 (omega^(n/2) = -1)
    s = a
    while s.shape[0] > 1:  # pairwise tree sum on the device field adder
        s = engine.test_field_op(1, 2, s[0::2].copy(), s[1::2].copy())
    assert (fa[0] == s[0]).all()
    assert (fb[n // 2] == orc.field_op("fr", 5, fa[n // 2:n // 2 + 1])[0]).all()


@pytest.mark.gpu
def test_gpu_cppoly_prove(engine, orc, golden):
    g = golden("fr_vectors")
    for d in g["prove_dims"]:
        d = int(d)
        P, v, r = g[f"prove_bases_{d}"], g[f"prove_v_{d}"], g[f"prove_r_{d}"]
        key = engine.CommitmentKey("g1", P)
        try:
            wit, ev = engine.cppoly_prove(key, v, r)
            assert (wit == g[f"prove_witness_{d}"]).all(), d
            assert (ev == orc.fr_eval_mle(v, r)).all()
        finally:
            key.close()
    # a size where the big levels run the multi-kernel pipeline, on a plain and on a precomputed key
    d = 14
    n = 1 << d
    from tests import inputs
    P, _ = inputs.bases(orc, "g1", n // 2, seed=91, affine=False)
    v = orc.sha512_rng_fr(92, n)
    r = orc.sha512_rng_fr(93, d)
    want = orc.cppoly_prove_g1(P, v, r)
    key = engine.CommitmentKey("g1", P)
    try:
        wit, _ = engine.cppoly_prove(key, v, r)
        assert (wit == want).all()
        key.precompute(9)
        engine.set_tuning_ex("use_precomputed", 2)
        wit, _ = engine.cppoly_prove(key, v, r)
        assert (wit == want).all()
    finally:
        engine.set_tuning_ex("use_precomputed", 1)
        key.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_geo,n_tail", [(1, 0), (31, 1), (32, 0), (33, 2), (4096 + 7, 1), (1 << 16, 1)])
def test_gpu_scale_inv_geometric(engine, orc, n_geo, n_tail):
    """b200_fr_scale_inv_geometric = the divisions of step_radix2_domain::divide_by_Z_on_coset
    (step_radix2_domain.tcc:213-241): P[i] * (c1 ratio^i - c0)^-1, checked with Python integers (the inverse of a
    field element is unique); run lengths on both sides of the per-thread batch of 32."""
    from oracle.binding import mont_to_ints
    n = n_geo + n_tail
    P = orc.sha512_rng_fr(2100 + n_geo, n)
    c = orc.sha512_rng_fr(2200 + n_geo, 4)
    got = engine.scale_inv_geometric(P, n_geo, c[0], c[1], c[2], c[3] if n_tail else None)
    Pi, (c1, ratio, c0, tail) = mont_to_ints(P, R_ORDER), mont_to_ints(c, R_ORDER)
    want, t = [], c1
    for i in range(n_geo):
        want.append(Pi[i] * pow((t - c0) % R_ORDER, -1, R_ORDER) % R_ORDER)
        t = t * ratio % R_ORDER
    for i in range(n_tail):
        want.append(Pi[n_geo + i] * tail % R_ORDER)
    assert (got == ints_to_mont(want, R_ORDER)).all()


@pytest.mark.gpu
def test_gpu_fr_argument_errors(engine):
    import legosnark_b200 as lb
    a = np.zeros((8, 4), dtype=np.uint64)
    with pytest.raises(lb.B200Error):
        engine.fr_fft(a, 2, None)          # coset transform without a shift
    with pytest.raises(ValueError):
        engine.fr_fft(a[:6], 0)            # libfqfft: DomainSizeException
    with pytest.raises(ValueError):
        engine.evalMLE(a, a[:2])           # assert(N == 1 << d)
