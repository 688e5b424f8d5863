"""Sum-check dynamic-programming tables (LS = /root/reference/src): DPBeta::compute_eq_tbl
(LS/prototools/mle.h:93-105), DPMatrixMle's constructor (mle.h:241-259) and the round polynomials of CPSumcheck::prove's
loop (LS/gadgets/sumcheck.cc:56-70: make_new_h_poly, LS/gadgets/sumcheck.h:85-106, + DPMle::pushRandomness).

CPU: the plain-C restatement against fixtures the reference produced (tests/golden/sumcheck.npz, tools/make_golden_sumcheck.py)
and against the reference live when oracle/_ref is present.  GPU: the C-ABI entry points against the same fixtures and the
restatement, and at 2^20 through properties the definitions give."""
import numpy as np
import pytest

from oracle.binding import R_ORDER, MONT_R, ints_to_mont, mont_to_ints


def _poly_times_beta(S, rho_j, pre=1):
    """(e0 + e1 x) * pre * S(x) with eqbit_poly(rho) = (1 - rho) + (2 rho - 1) x (mle.cc:25-31): 4 coefficients."""
    s = mont_to_ints(S, R_ORDER)
    rj = mont_to_ints(rho_j.reshape(1, 4), R_ORDER)[0]
    e0, e1 = (1 - rj) * pre % R_ORDER, (2 * rj - 1) * pre % R_ORDER
    out = [0, 0, 0, 0]
    for k in range(3):
        out[k] = (out[k] + e0 * s[k]) % R_ORDER
        out[k + 1] = (out[k + 1] + e1 * s[k]) % R_ORDER
    return ints_to_mont(out, R_ORDER)


def _cases(g):
    return [int(d) for d in g["dims"]], [int(d) for d in g["matrix_dims"]]


# ---------------------------------------------------------------- CPU: oracle pinned
def test_oracle_vs_reference_fixtures(orc, golden):
    g = golden("sumcheck")
    dims, mdims = _cases(g)
    for d in dims:
        assert (orc.fr_eq_table(g[f"rho_{d}"]) == g[f"eq_{d}"]).all(), d
        assert (orc.fr_eq_table(g[f"r_{d}"]) == g[f"eq_r_{d}"]).all(), d
        assert (orc.fr_sumcheck_rounds(g[f"a_{d}"], g[f"b_{d}"], g[f"r_{d}"]) == g[f"h_dummy_{d}"]).all(), d
        if d >= 2:  # round 0 with a real beta: the restated weighted sum times eqbit_poly(rho[0]) (beta_pre = 1)
            S = orc.fr_sumcheck_round(g[f"a_{d}"], g[f"b_{d}"], g[f"beta_suffix_{d}"])
            assert (_poly_times_beta(S, g[f"rho_{d}"][0]) == g[f"h_beta_{d}"][0]).all(), d
    for d in mdims:
        assert (orc.fr_matrix_mle(g[f"A_{d}"], g[f"mrho_{d}"]) == g[f"matrix_mle_{d}"]).all(), d


def test_oracle_vs_reference_live(orc, ref):
    for d in (1, 4, 7):
        rho, r = orc.sha512_rng_fr(50 + d, d), orc.sha512_rng_fr(60 + d, d)
        a, b = orc.sha512_rng_fr(70 + d, 1 << d), orc.sha512_rng_fr(80 + d, 1 << d)
        assert (orc.fr_eq_table(rho) == ref.fr_eq_table(rho)).all()
        assert (orc.fr_sumcheck_rounds(a, b, r) == ref.fr_sumcheck_rounds(a, b, r)).all()
    for d in (3, 5):
        A, rho = orc.sha512_rng_fr(90 + d, 1 << (2 * d)), orc.sha512_rng_fr(95 + d, d)
        assert (orc.fr_matrix_mle(A, rho) == ref.fr_matrix_mle(A, rho)).all()


def test_oracle_step_domain_vs_reference_fixtures(orc, golden):
    """libfqfft's step_radix2_domain (the domain of 2^k + 2^r constraints: BASELINE.json configs[3] has 2^21 + 1)."""
    g = golden("sumcheck")
    g5 = ints_to_mont([5], R_ORDER)
    for lb, ls in g["step_shapes"]:
        lb, ls = int(lb), int(ls)
        a = g[f"step_a_{lb}_{ls}"]
        for mode in range(4):
            assert (orc.fr_step_fft(a, lb, ls, mode, g5) == g[f"step_{lb}_{ls}_m{mode}"]).all(), (lb, ls, mode)


def _step_divide_z_expected(a, lb, ls):
    """step_radix2_domain::divide_by_Z_on_coset (step_radix2_domain.tcc:213-241) with Python integers."""
    r = R_ORDER
    big, small = 1 << lb, 1 << ls
    rou = 19103219067921713944291392827692070036145651957329286315305642004821462161904  # Fr::root_of_unity, order 2^28
    omega = pow(rou, 1 << (28 - (lb + 1)), r)
    coset = 5
    Z0 = (pow(coset, big, r) - 1) % r
    c1, c0, ratio = pow(coset, small, r) * Z0 % r, pow(omega, small, r) * Z0 % r, pow(omega, 2 * small, r)
    Z1 = (pow(coset * omega, big, r) - 1) * (pow(coset * omega, small, r) - pow(omega, small, r)) % r
    ai = mont_to_ints(a, r)
    out, elt = [], 1
    for i in range(big):
        out.append(ai[i] * pow((c1 * elt - c0) % r, -1, r) % r)
        elt = elt * ratio % r
    z1i = pow(Z1, -1, r)
    out += [ai[big + i] * z1i % r for i in range(small)]
    return ints_to_mont(out, r), (c1, ratio, c0, z1i)


def test_step_divide_z_formula_vs_reference_fixtures(golden):
    g = golden("sumcheck")
    for lb, ls in g["step_shapes"]:
        lb, ls = int(lb), int(ls)
        want, _ = _step_divide_z_expected(g[f"step_a_{lb}_{ls}"], lb, ls)
        assert (want == g[f"step_{lb}_{ls}_divz"]).all(), (lb, ls)


# ---------------------------------------------------------------- GPU: parity through the C-ABI
@pytest.mark.gpu
def test_gpu_step_domain_fixtures(engine, golden):
    g = golden("sumcheck")
    g5 = ints_to_mont([5], R_ORDER)
    for lb, ls in g["step_shapes"]:
        lb, ls = int(lb), int(ls)
        a = g[f"step_a_{lb}_{ls}"]
        for mode in range(4):
            assert (engine.fr_step_fft(a, lb, ls, mode, g5) == g[f"step_{lb}_{ls}_m{mode}"]).all(), (lb, ls, mode)
        _, (c1, ratio, c0, z1i) = _step_divide_z_expected(a, lb, ls)
        c = ints_to_mont([c1, ratio, c0, z1i], R_ORDER)
        got = engine.scale_inv_geometric(a, 1 << lb, c[0], c[1], c[2], c[3])
        assert (got == g[f"step_{lb}_{ls}_divz"]).all(), (lb, ls, "divide_by_Z_on_coset")


@pytest.mark.gpu
@pytest.mark.parametrize("lb,ls", [(1, 0), (10, 0), (11, 10), (12, 5), (14, 0), (17, 3)])
def test_gpu_step_domain_vs_oracle(engine, orc, lb, ls):
    a = orc.sha512_rng_fr(3300 + lb, (1 << lb) + (1 << ls))
    g = orc.sha512_rng_fr(3400 + lb, 1)
    for mode in range(4):
        assert (engine.fr_step_fft(a, lb, ls, mode, g) == orc.fr_step_fft(a, lb, ls, mode, g)).all(), (lb, ls, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("lb,ls", [(3, 0), (6, 2), (10, 0), (12, 11), (13, None), (5, None)])
def test_gpu_qap_h_coefficients(engine, orc, lb, ls):
    """b200_qap_h_coefficients = the vector part of r1cs_to_qap_witness_map (r1cs_to_qap.tcc:232-311, d1 = d2 = d3 = 0):
    iFFT x 3, cosetFFT x 3, A B - C, divide_by_Z_on_coset, icosetFFT, composed here from the pinned restatements
    (orc_fr_step_fft / orc_fr_fft, the reference-checked division formula) on step domains (2^r = 1 and > 1) and basic ones."""
    r = R_ORDER
    m = (1 << lb) + (0 if ls is None else 1 << ls)
    a, b, c = (orc.sha512_rng_fr(4400 + 10 * lb + k, m) for k in range(3))
    g5 = ints_to_mont([5], r)

    def fwd(v, mode):
        return orc.fr_fft(v, mode, g5) if ls is None else orc.fr_step_fft(v, lb, ls, mode, g5)
    A, B, C = (fwd(fwd(v, 1), 2) for v in (a, b, c))
    T = orc.field_op("fr", 3, orc.field_op("fr", 0, A, B), C)  # A * B - C
    if ls is None:
        zinv = pow((pow(5, m, r) - 1) % r, -1, r)
        T = ints_to_mont([x * zinv % r for x in mont_to_ints(T, r)], r)
        div = ints_to_mont([zinv], r)
    else:
        T, (c1, ratio, c0, z1i) = _step_divide_z_expected(T, lb, ls)
        div = ints_to_mont([c1, ratio, c0, z1i], r)
    want = fwd(T, 3)
    got = engine.qap_h_coefficients(a, b, c, lb, ls, g5, div)
    assert (got == want).all()


@pytest.mark.gpu
def test_gpu_step_domain_round_trip_at_scale(engine):
    """2^21 + 1 points (the 128 x 128 matrix product): iFFT(FFT(a)) = a and icosetFFT(cosetFFT(a)) = a."""
    lb, ls = 21, 0
    m = (1 << lb) + 1
    rng = np.random.default_rng(11)
    a = rng.integers(0, 1 << 64, size=(m, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    g5 = ints_to_mont([5], R_ORDER)
    assert (engine.fr_step_fft(engine.fr_step_fft(a, lb, ls, 0), lb, ls, 1) == a).all()
    assert (engine.fr_step_fft(engine.fr_step_fft(a, lb, ls, 2, g5), lb, ls, 3, g5) == a).all()

@pytest.mark.gpu
def test_gpu_sumcheck_fixtures(engine, golden):
    g = golden("sumcheck")
    dims, mdims = _cases(g)
    for d in dims:
        a, b, r, rho = g[f"a_{d}"], g[f"b_{d}"], g[f"r_{d}"], g[f"rho_{d}"]
        assert (engine.compute_eq_tbl(rho) == g[f"eq_{d}"]).all(), d
        assert (engine.compute_eq_tbl(r) == g[f"eq_r_{d}"]).all(), d
        assert (engine.sumcheck_rounds(a, b, r) == g[f"h_dummy_{d}"]).all(), d
        assert (engine.sumcheck_round(a, b) == g[f"h_dummy_{d}"][0]).all(), d
        if d >= 2:
            S = engine.sumcheck_round(a, b, g[f"beta_suffix_{d}"])
            assert (_poly_times_beta(S, rho[0]) == g[f"h_beta_{d}"][0]).all(), d
    for d in mdims:
        assert (engine.matrix_mle(g[f"A_{d}"], g[f"mrho_{d}"]) == g[f"matrix_mle_{d}"]).all(), d


@pytest.mark.gpu
@pytest.mark.parametrize("d", [1, 6, 9, 10, 13, 16])
def test_gpu_sumcheck_vs_oracle(engine, orc, d):
    rho, r = orc.sha512_rng_fr(150 + d, d), orc.sha512_rng_fr(160 + d, d)
    a, b = orc.sha512_rng_fr(170 + d, 1 << d), orc.sha512_rng_fr(180 + d, 1 << d)
    w = orc.sha512_rng_fr(190 + d, 1 << (d - 1)) if d >= 1 else None
    assert (engine.compute_eq_tbl(rho) == orc.fr_eq_table(rho)).all()
    assert (engine.sumcheck_rounds(a, b, r) == orc.fr_sumcheck_rounds(a, b, r)).all()
    assert (engine.sumcheck_round(a, b, w) == orc.fr_sumcheck_round(a, b, w)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("d", [1, 5, 7, 9])
def test_gpu_matrix_mle_vs_oracle(engine, orc, d):
    """d = 7 is BASELINE.json configs[3] (128 x 128 matrices); d = 9 has 2^18 entries and several row slices per column."""
    A, rho = orc.sha512_rng_fr(250 + d, 1 << (2 * d)), orc.sha512_rng_fr(260 + d, d)
    assert (engine.matrix_mle(A, rho) == orc.fr_matrix_mle(A, rho)).all()


@pytest.mark.gpu
def test_gpu_sumcheck_properties_at_scale(engine, orc):
    """2^20-entry tables: h_0(0) + h_0(1) = sum_p a[p] b[p] (the claim the first round polynomial must meet), and the last
    round's h(r) = evalMLE(a, r') * evalMLE(b, r') chain, checked through the engine's own evalMLE on the bound tables."""
    d = 20
    n = 1 << d
    rng = np.random.default_rng(7)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    b = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    b[:, 3] &= np.uint64((1 << 60) - 1)
    r = orc.sha512_rng_fr(321, d)
    h = engine.sumcheck_rounds(a, b, r)
    # h_i(0) + h_i(1) = 2 c0 + c1 + c2 must equal h_{i-1}(r_{i-1}); for i = 0 the inner product of the tables
    def at(coeffs, x):
        c = mont_to_ints(coeffs, R_ORDER)
        return (c[0] + c[1] * x + c[2] * x * x) % R_ORDER
    rr = mont_to_ints(r, R_ORDER)
    prod = engine.test_field_op(1, 0, a, b)  # Montgomery product a[p] * b[p] on the device
    s = prod
    while s.shape[0] > 1:
        s = engine.test_field_op(1, 2, s[0::2].copy(), s[1::2].copy())
    claim = mont_to_ints(s[:1], R_ORDER)[0]
    for i in range(d):
        assert (at(h[i], 0) + at(h[i], 1)) % R_ORDER == claim, i
        claim = at(h[i], rr[i])
