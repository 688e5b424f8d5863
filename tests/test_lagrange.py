"""evaluate_all_lagrange_polynomials of libfqfft's radix-2 and step domains (the Groth16 generator's side of the domain:
r1cs_to_qap_instance_map_with_evaluation, SNK/reductions/r1cs_to_qap/r1cs_to_qap.tcc:127-190).  The reference spends one field
inversion per domain point (FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.tcc:228-233, step_radix2_domain.tcc:172-176);
the engine batches them on the device (b200_fr_geometric_quotients).  Fixtures: tests/golden/lagrange.npz, written by the
reference (tools/make_golden_lagrange.py)."""
import numpy as np
import pytest

from oracle.binding import R_ORDER, ints_to_mont, mont_to_ints


def _cases(g):
    for lb, ls in g["shapes"]:
        lb, ls = int(lb), int(ls)
        for tk, uk in (("t", "u"), ("tin", "uin"), ("tsm", "usm")):
            if f"{tk}_{lb}_{ls}" in g:
                yield lb, (None if ls < 0 else ls), g[f"{tk}_{lb}_{ls}"], g[f"{uk}_{lb}_{ls}"], tk


def test_oracle_lagrange_vs_reference_fixtures(orc, golden):
    g = golden("lagrange")
    n = 0
    for lb, ls, t, want, kind in _cases(g):
        assert (orc.fr_lagrange(lb, ls, t) == want).all(), (lb, ls, kind)
        n += 1
    assert n == 29


def test_oracle_lagrange_vs_reference_live(orc, ref):
    for lb, ls in [(1, None), (4, None), (10, None), (2, 1), (5, 3), (10, 0), (10, 9)]:
        t = ref.sha512_rng_fr(500 + lb, 1)
        assert (orc.fr_lagrange(lb, ls, t) == ref.fr_lagrange(lb, ls, t)).all(), (lb, ls)


def test_lagrange_fixtures_interpolate(golden):
    """What the vector means, with Python integers: sum_i u[i] * P(x_i) = P(t) for a polynomial of degree < m."""
    g = golden("lagrange")
    r = R_ORDER
    rou = 19103219067921713944291392827692070036145651957329286315305642004821462161904
    for lb, ls in [(5, -1), (4, 2), (3, 0)]:
        u = mont_to_ints(g[f"u_{lb}_{ls}"], r)
        t = mont_to_ints(g[f"t_{lb}_{ls}"], r)[0]
        if ls < 0:
            w = pow(rou, 1 << (28 - lb), r)
            xs = [pow(w, i, r) for i in range(1 << lb)]
        else:  # step_radix2_domain::get_domain_element (step_radix2_domain.tcc:189-199)
            omega = pow(rou, 1 << (28 - (lb + 1)), r)
            so = pow(rou, 1 << (28 - ls), r) if ls else 1
            xs = [pow(omega * omega % r, i, r) for i in range(1 << lb)] + [omega * pow(so, i, r) % r for i in range(1 << ls)]
        coeffs = [(7 * i * i + 3 * i + 11) % r for i in range(len(xs))]
        P = lambda x: sum(c * pow(x, k, r) for k, c in enumerate(coeffs)) % r  # noqa: E731
        assert sum(ui * P(x) for ui, x in zip(u, xs)) % r == P(t), (lb, ls)


# ---------------------------------------------------------------- GPU: parity through the C-ABI
@pytest.mark.gpu
def test_gpu_lagrange_fixtures(engine, golden):
    g = golden("lagrange")
    for lb, ls, t, want, kind in _cases(g):
        got = engine.evaluate_all_lagrange_polynomials(lb, ls, t)
        assert got.shape == want.shape and (got == want).all(), (lb, ls, kind)


@pytest.mark.gpu
@pytest.mark.parametrize("lb,ls", [(7, None), (13, None), (16, None), (13, 0), (14, 9), (16, 15)])
def test_gpu_lagrange_vs_oracle(engine, orc, lb, ls):
    t = orc.sha512_rng_fr(4100 + lb, 1)
    assert (engine.evaluate_all_lagrange_polynomials(lb, ls, t) == orc.fr_lagrange(lb, ls, t)).all()


@pytest.mark.gpu
def test_gpu_geometric_quotients_with_input_and_in_place_sizes(engine, orc):
    """in != NULL (the step domain's pass over the unit vector) and sizes around the per-thread run of 32 / the block of 4096."""
    r = R_ORDER
    for n in (1, 31, 32, 33, 4095, 4096, 4097, 20000):
        vals = [int(x) % r for x in np.random.default_rng(n).integers(1, 1 << 62, size=7)]
        a0, ar, c1, ra, c0, d1, rb = vals
        d0 = 12345
        inp = orc.sha512_rng_fr(4200 + n, n)
        ii = mont_to_ints(inp, r)
        for nf in (1, 2):
            rows = [a0, ar, c1, ra, c0] + ([d1, rb, d0] if nf == 2 else [])
            want = []
            for i in range(min(n, 40)):  # spot-check the head exactly, the rest through the oracle-free identity below
                den = (c1 * pow(ra, i, r) - c0) % r
                if nf == 2:
                    den = den * ((d1 * pow(rb, i, r) - d0) % r) % r
                want.append(ii[i] * a0 * pow(ar, i, r) % r * pow(den, -1, r) % r)
            got = engine.geometric_quotients(n, ints_to_mont(rows, r), inp)
            assert mont_to_ints(got[: len(want)], r) == want, (n, nf)
            ones = engine.geometric_quotients(n, ints_to_mont(rows, r))
            gi, oi = mont_to_ints(got[-3:], r), mont_to_ints(ones[-3:], r)
            assert [o * x % r for o, x in zip(oi, ii[-3:])] == gi, (n, nf, "tail")


@pytest.mark.gpu
def test_gpu_lagrange_config4_domain_interpolates(engine):
    """The domain of BASELINE.json configs[3] (128 x 128 matrix product: 2^21 + 1 points): sum_i u[i] = 1 and
    sum_i u[i] x_i = t (the Lagrange basis reproduces the polynomials 1 and x), with Python integers."""
    lb, ls, r = 21, 0, R_ORDER
    t_int = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % r
    u = engine.evaluate_all_lagrange_polynomials(lb, ls, ints_to_mont([t_int], r))
    assert u.shape[0] == (1 << lb) + 1
    rou = 19103219067921713944291392827692070036145651957329286315305642004821462161904
    omega = pow(rou, 1 << (28 - (lb + 1)), r)
    ui = mont_to_ints(u, r)
    assert sum(ui) % r == 1
    s, x, bo = 0, 1, omega * omega % r
    for i in range(1 << lb):
        s += ui[i] * x
        x = x * bo % r
    s += ui[1 << lb] * omega
    assert s % r == t_int
