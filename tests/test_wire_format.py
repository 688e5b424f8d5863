"""Point compression of the reference's stream operators: operator<< / operator>> of
alt_bn128_G1/G2 (alt_bn128_g1.cpp:404-459, alt_bn128_g2.cpp:414-475) and bn128_G1/G2 (bn128_g1.cpp:344-463,
bn128_g2.cpp:374-470).  CPU: the restatement against fixtures written by the reference's own operators
(tools/make_golden_wire.py) and against the reference live.  GPU (-m gpu): the CUDA path through the C-ABI against
the same fixtures and the oracle, round trips at key sizes, and the error path for X not on the curve."""
import numpy as np
import pytest

from tests import inputs

CASES = [("g1", 0), ("g1", 2), ("g2", 0), ("g2", 2)]


def _same_points(a, b, grp):
    """Equal as (X, Y, Z) images, treating every zero alike (alt_bn128 reads a zero back as (0,1,0), bn128 as (1,1,0))."""
    L = a.shape[1]
    za = (a[:, 2 * L // 3:] == 0).all(axis=1)
    zb = (b[:, 2 * L // 3:] == 0).all(axis=1)
    return (za == zb).all() and (a[~za] == b[~zb]).all()


@pytest.mark.parametrize("grp,fl", CASES)
def test_oracle_wire_golden(orc, golden, grp, fl):
    g = golden("wire")
    P, x, flags = g[f"{grp}_f{fl}_points"], g[f"{grp}_f{fl}_x"], g[f"{grp}_f{fl}_flags"]
    ox, oflags = orc.compress(grp, P, fl)
    assert (ox == x).all() and (oflags == flags).all()
    assert _same_points(orc.decompress(grp, x, flags, fl), g[f"{grp}_f{fl}_read_back"], grp)


@pytest.mark.parametrize("grp,fl", CASES)
def test_oracle_wire_vs_reference_live(orc, ref, grp, fl):
    P, _ = inputs.bases(orc, grp, 24, seed=123 + fl, affine=False)
    P[0] = inputs.zero_point(grp, curve=1 if fl == 2 else 0)
    rx, rflags = ref.compress(grp, P, fl)
    ox, oflags = orc.compress(grp, P, fl)
    assert (ox == rx).all() and (oflags == rflags).all()
    assert _same_points(orc.decompress(grp, rx, rflags, fl), ref.decompress(grp, rx, rflags, fl), grp)


def test_oracle_montgomery_output_flavour(orc):
    """flavour 1 (-DMONTGOMERY_OUTPUT): X is the Montgomery image, the Y bit still comes from as_bigint (fp.tcc operator<<)."""
    P, _ = inputs.bases(orc, "g1", 16, seed=5, affine=True)
    x0, f0 = orc.compress("g1", P, 0)
    x1, f1 = orc.compress("g1", P, 1)
    assert (f0 == f1).all() and (x1 == P[:, :4]).all() and (orc.fq_from_bigint(x0) == x1).all()
    assert (orc.decompress("g1", x1, f1, 1) == P).all()


@pytest.mark.gpu
@pytest.mark.parametrize("grp,fl", CASES)
def test_gpu_wire_golden(engine, golden, grp, fl):
    g = golden("wire")
    P, x, flags = g[f"{grp}_f{fl}_points"], g[f"{grp}_f{fl}_x"], g[f"{grp}_f{fl}_flags"]
    gx, gflags = engine.compress_points(grp, P, fl)
    assert (gx == x).all() and (gflags == flags).all()
    assert _same_points(engine.decompress_points(grp, x, flags, fl), g[f"{grp}_f{fl}_read_back"], grp)


@pytest.mark.gpu
@pytest.mark.parametrize("grp,n", [("g1", 5000), ("g2", 1500)])
def test_gpu_wire_vs_oracle_and_round_trip(engine, orc, grp, n):
    P, _ = inputs.bases(orc, grp, n, seed=321, affine=False)
    P[7] = inputs.zero_point(grp)
    aff = orc.batch_to_special(grp, P)
    for fl in (0, 1, 2):
        x, flags = engine.compress_points(grp, P, fl)
        ox, oflags = orc.compress(grp, P, fl)
        assert (x == ox).all() and (flags == oflags).all(), fl
        back = engine.decompress_points(grp, x, flags, fl)
        assert (back == aff).all(), fl  # (x, y, 1) / (0, 1, 0): batch_to_special's image
        assert (back == orc.decompress(grp, x, flags, fl)).all(), fl
    assert engine.compress_points(grp, P[:0], 0)[0].shape[0] == 0


@pytest.mark.gpu
def test_gpu_wire_key_size_round_trip(engine, orc):
    """A 2^18-point key made on the device, written and read back (size-independent property: identity)."""
    n = 1 << 18
    k = inputs.fr_uniform(orc, n, seed=9)
    table = engine.get_window_table("g1", 254, 0, orc.one("g1"), expected_scalars=n)
    P = engine.batch_exp(254, 0, table, k)
    table.close()
    x, flags = engine.compress_points("g1", P, 2)
    assert (engine.decompress_points("g1", x, flags, 2) == P).all()
    idx = np.r_[0:32, n - 32:n]
    ox, oflags = orc.compress("g1", P[idx], 2)
    assert (x[idx] == ox).all() and (flags[idx] == oflags).all()


@pytest.mark.gpu
def test_gpu_wire_not_on_curve(engine, orc):
    """X with X^3 + 3 a non-residue: the reference's sqrt would not terminate; the engine reports it."""
    import legosnark_b200 as lb
    x = np.zeros((64, 4), dtype=np.uint64)
    x[:, 0] = np.arange(64, dtype=np.uint64) + 1
    flags = np.zeros(64, dtype=np.uint8)
    pts, bad = engine.decompress_points("g1", x, flags, 0, report_bad=True)
    assert 0 < bad.sum() < 64
    good = bad == 0
    assert (orc.decompress("g1", x[good], flags[good], 0) == pts[good]).all()
    with pytest.raises(lb.B200Error):
        engine.decompress_points("g1", x, flags, 0)
