"""GPU: multi_exp / multi_exp_with_mixed_addition through the C-ABI, bit-exact (after
affine normalisation) against the reference fixtures, the oracle at sizes it finishes in
seconds, and the scalar-sum identity of SURVEY.md §8(c) at full benchmark sizes."""
import numpy as np
import pytest

from oracle.binding import R_ORDER, ints_to_mont, limbs_to_int, MONT_R
from tests import inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_golden_cases(engine, golden, grp):
    g = golden(f"msm_{grp}")
    for name in g["names"]:
        B, S, R = g[f"{name}__bases"], g[f"{name}__scalars"], g[f"{name}__result"]
        assert (engine.multi_exp(grp, B, S) == R).all(), (grp, name)
        assert (engine.multi_exp_with_mixed_addition(grp, B, S, chunks=8) == R).all(), (grp, name)


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_golden_cases_forced_geometry(engine, golden, grp):
    """Small windows / tiny task length force multi-task buckets, the warp combine and the
    segment weighting paths on the same fixtures."""
    g = golden(f"msm_{grp}")
    try:
        for c, L in ((2, 1), (3, 2), (5, 3), (8, 4), (13, 7)):
            engine.set_tuning(c, L)
            for name in g["names"]:
                B, S, R = g[f"{name}__bases"], g[f"{name}__scalars"], g[f"{name}__result"]
                assert (engine.multi_exp(grp, B, S) == R).all(), (grp, name, c, L)
    finally:
        engine.set_tuning(0, 0)


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_pipelined_upload_chunks(engine, orc, golden, grp):
    """Host-buffer MSMs cut into upload chunks (H2D of chunk j+1 under the accumulation of chunk j,
    bucket sums folded into the dense array): every edge-case fixture and a uniform case, for
    chunk counts that do and do not divide n."""
    g = golden(f"msm_{grp}")
    n = 9001 if grp == "g1" else 2050
    P, _ = inputs.bases(orc, grp, n, seed=41, affine=False)
    s = inputs.fr_uniform(orc, n, seed=42)
    want = orc.msm(grp, P, s)
    try:
        # dense_direct = 1 (default): whole buckets go from k_accumulate straight into the array that lives across the chunks and
        # only split buckets are folded; 0: every chunk folds every bucket
        # overlap_sort = 1 (default): chunk j + 1 is sorted on a second stream (two sets of sort outputs) under the accumulation
        # of chunk j; (1, 0, *): the round-1 counting sort feeds the same lists
        for dense_direct, part_sort, overlap in ((1, 1, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0), (0, 0, 0)):
            engine.set_tuning_ex("dense_direct", dense_direct)
            engine.set_tuning_ex("partition_sort", part_sort)
            engine.set_tuning_ex("overlap_sort", overlap)
            for chunks, c, L in ((2, 0, 0), (3, 7, 4), (4, 10, 0), (7, 4, 2), (16, 12, 0)):
                engine.set_tuning(c, L)
                engine.set_pipeline_chunks(chunks)
                assert (engine.multi_exp(grp, P, s) == want).all(), (grp, chunks, c, L, dense_direct, part_sort, overlap)
                if c == 0:
                    continue  # fixtures are small: they take the single-kernel path unless the geometry is forced
                for name in g["names"]:
                    B, S, R = g[f"{name}__bases"], g[f"{name}__scalars"], g[f"{name}__result"]
                    assert (engine.multi_exp(grp, B, S) == R).all(), (grp, name, chunks, c, L, dense_direct)
    finally:
        engine.set_tuning_ex("dense_direct", 1)
        engine.set_tuning_ex("partition_sort", 1)
        engine.set_tuning_ex("overlap_sort", 1)
        engine.set_tuning(0, 0)
        engine.set_pipeline_chunks(0)


@pytest.mark.parametrize("grp,n", [("g1", 1 << 12), ("g1", (1 << 14) + 7), ("g2", 1 << 10)])
def test_uniform_vs_oracle(engine, orc, grp, n):
    P, k = inputs.bases(orc, grp, n, seed=301)
    s = inputs.fr_uniform(orc, n, seed=302)
    want = orc.msm(grp, P, s, chunks=orc.max_threads())
    assert (engine.multi_exp(grp, P, s) == want).all()
    assert (inputs.scalar_sum_check(orc, grp, k, s) == want).all()


@pytest.mark.parametrize("grp,n", [("g1", 3000), ("g2", 700)])
def test_distributions_vs_oracle(engine, orc, grp, n):
    P, _ = inputs.bases(orc, grp, n, seed=311)
    PJ, _ = inputs.bases(orc, grp, n, seed=312, affine=False)  # callers pass Z != 1 (cplink.cc:56)
    cases = {
        "jacobian": (PJ, inputs.fr_uniform(orc, n, seed=313)),
        "32bit": (P, inputs.fr_small(n, 32)),
        "zero_one_heavy": (P, inputs.fr_zero_one_heavy(orc, n)),
        "all_equal_bases": (np.tile(PJ[3], (n, 1)), inputs.fr_uniform(orc, n, seed=314)),
        "all_ones": (P, inputs.fr_const(n, 1)),
        "hot_bucket": (P, inputs.fr_const(n, 0xabcdef)),
    }
    for name, (B, S) in cases.items():
        assert (engine.multi_exp(grp, B, S) == orc.msm(grp, B, S, chunks=orc.max_threads(), variant=1)).all(), name


def test_cplink_shape_prove(engine, orc):
    """cplink prove: one G1 MSM, n = 1026, Jacobian bases (LS/gadgets/subspace.cc:78-85)."""
    n = 1026
    P, _ = inputs.bases(orc, "g1", n, seed=321, affine=False)
    w = inputs.fr_uniform(orc, n, seed=322)
    assert (engine.multi_exp_with_mixed_addition("g1", P, w) == orc.msm("g1", P, w, variant=1)).all()


@pytest.mark.parametrize("grp,n", [("g1", 5000), ("g2", 600)])
def test_pinned_key_prefixes(engine, orc, grp, n):
    """CPPoly::prove runs MSMs over prefixes of one key (LS/gadgets/poly.h:77-88)."""
    P, _ = inputs.bases(orc, grp, n, seed=331, affine=False)
    key = engine.CommitmentKey(grp, P)
    try:
        for m in (n, n // 2, n // 4, 1, 0):
            s = inputs.fr_uniform(orc, m, seed=332 + m)
            assert (key.multi_exp(s) == orc.msm(grp, P[:m], s, chunks=orc.max_threads())).all(), m
        s = inputs.fr_uniform(orc, 100, seed=340)
        assert (key.multi_exp(s, offset=37) == orc.msm(grp, P[37:137], s)).all()
    finally:
        key.close()


@pytest.mark.parametrize("grp,n", [("g1", 6000), ("g2", 4500)])
def test_precomputed_key(engine, orc, golden, grp, n):
    """b200_key_precompute_*: the key also holds 2^(c k) P_i, all windows share one bucket set.
    Same group element as the plain key and the oracle for several window sizes, on prefixes /
    offset sub-ranges, with zero and repeated bases, P/-P pairs and skewed scalars."""
    P, _ = inputs.bases(orc, grp, n, seed=351, affine=False)
    P[5] = inputs.zero_point(grp)  # a zero base
    P[7] = P[6]                   # repeated bases -> doubling branch inside a bucket
    P[9] = inputs.negate(orc, grp, P[8:9])[0]  # opposite bases
    key = engine.CommitmentKey(grp, P)
    try:
        engine.set_tuning_ex("use_precomputed", 2)  # always, also where the cost model prefers the plain path
        for c in (4, 7, 11, 0):
            key.precompute(c)
            for m, off in ((n, 0), (n - 1, 1), (4100, 333)):
                for s in (inputs.fr_uniform(orc, m, seed=352 + m), inputs.fr_zero_one_heavy(orc, m)):
                    s[3:12] = s[3]  # equal scalars on the special bases
                    want = orc.msm(grp, P[off:off + m], s, chunks=orc.max_threads())
                    got = key.multi_exp(s, offset=off)
                    assert (got == want).all(), (grp, c, m, off)
                    engine.set_tuning_ex("host_horner", 0)  # the device weights and sums the per-job results itself
                    try:
                        assert (key.multi_exp(s, offset=off) == want).all(), (grp, c, m, off, "device horner")
                    finally:
                        engine.set_tuning_ex("host_horner", 1)
            if c:
                st = engine.last_stats()
                assert st["window_bits"] == c, "precomputed path was not taken"
        # short sub-ranges fall back to the plain / single-kernel paths
        s = inputs.fr_uniform(orc, 100, seed=360)
        assert (key.multi_exp(s, offset=37) == orc.msm(grp, P[37:137], s)).all()
        # switching the levels off gives the plain path on the same key
        engine.set_tuning_ex("use_precomputed", 0)
        s = inputs.fr_uniform(orc, n, seed=361)
        assert (key.multi_exp(s) == orc.msm(grp, P, s, chunks=orc.max_threads())).all()
    finally:
        engine.set_tuning_ex("use_precomputed", 1)
        key.close()


@pytest.mark.parametrize("grp,n", [("g1", 7001), ("g2", 4600)])
def test_resident_key_chunked_scalars(engine, orc, grp, n):
    """A commitment under a resident key with HOST scalars uploads them in index chunks from 2^19 points on (chunk j + 1 is
    uploaded and sorted under the accumulation of chunk j; `pinned_chunks` forces the chunk count here): plain and precomputed
    keys, sub-ranges at an offset, zero / repeated / opposite bases, skewed scalars (the ones filter across chunks), the
    per-job host Horner from the dense array and the device-side weighting."""
    P, _ = inputs.bases(orc, grp, n, seed=881, affine=False)
    P[5] = inputs.zero_point(grp)
    P[7] = P[6]
    P[9] = inputs.negate(orc, grp, P[8:9])[0]
    key = engine.CommitmentKey(grp, P)
    try:
        for pre_c in (0, 6, 11):
            if pre_c:
                key.precompute(pre_c)
                engine.set_tuning_ex("use_precomputed", 2)
            for chunks in (2, 3, 5):
                engine.set_tuning_ex("pinned_chunks", chunks)
                for m, off in ((n, 0), (n - 3, 2), (4200, 311)):
                    for s in (inputs.fr_uniform(orc, m, seed=882 + m), inputs.fr_zero_one_heavy(orc, m, seed=883)):
                        s[3:12] = s[3]
                        want = orc.msm(grp, P[off:off + m], s, chunks=orc.max_threads(), variant=1)
                        assert (key.multi_exp(s, offset=off) == want).all(), (grp, pre_c, chunks, m, off)
                        if pre_c:
                            engine.set_tuning_ex("host_horner", 0)
                            try:
                                assert (key.multi_exp(s, offset=off) == want).all(), (grp, pre_c, chunks, m, off, "device horner")
                            finally:
                                engine.set_tuning_ex("host_horner", 1)
    finally:
        engine.set_tuning_ex("pinned_chunks", 0)
        engine.set_tuning_ex("use_precomputed", 1)
        key.close()


@pytest.mark.parametrize("grp,n", [("g1", 9000), ("g2", 5000)])
def test_ones_filter(engine, orc, grp, n):
    """Scalars equal to one are summed directly (multi_exp_with_mixed_addition's special addition,
    multiexp.tcc:455-487) instead of going through the sort: same element with the filter on and off, with
    zero bases and repeated bases among the ones, all-one scalars, and through the pipelined upload."""
    P, _ = inputs.bases(orc, grp, n, seed=371, affine=False)
    P[3] = inputs.zero_point(grp)
    P[11] = P[10]
    one = ints_to_mont([1], R_ORDER)[0]
    cases = {"heavy": inputs.fr_zero_one_heavy(orc, n, seed=5), "all_one": np.tile(one, (n, 1)), "uniform": inputs.fr_uniform(orc, n, seed=372)}
    cases["heavy"][0:16] = one
    cases["uniform"][7] = one
    key = engine.CommitmentKey(grp, P)
    try:
        for name, s in cases.items():
            want = orc.msm(grp, P, s, chunks=orc.max_threads(), variant=1)
            for ones in (1, 0):
                engine.set_tuning_ex("ones_filter", ones)
                assert (key.multi_exp(s) == want).all(), (name, ones, "resident")
                assert (engine.multi_exp_with_mixed_addition(grp, P, s) == want).all(), (name, ones, "host")
            engine.set_tuning_ex("ones_filter", 1)
            engine.set_pipeline_chunks(3)
            assert (engine.multi_exp(grp, P, s) == want).all(), (name, "chunks")
            engine.set_pipeline_chunks(0)
            m = n // 2
            assert (key.multi_exp(s[:m], offset=100) == orc.msm(grp, P[100:100 + m], s[:m], chunks=orc.max_threads())).all()
    finally:
        engine.set_tuning_ex("ones_filter", 1)
        engine.set_pipeline_chunks(0)
        key.close()


def _scalar_sum_expected(orc, grp, k, s):
    """(sum s_i k_i mod r) * G, with the big-integer sum vectorised over 64-bit limbs."""
    rinv = pow(MONT_R, -1, R_ORDER)

    def to_ints(a):
        a = np.asarray(a, dtype=np.uint64)
        return [((int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192) * rinv) % R_ORDER for r in a]

    tot = sum(x * y for x, y in zip(to_ints(k), to_ints(s))) % R_ORDER
    return orc.scalar_mul(grp, orc.one(grp), ints_to_mont([tot], R_ORDER), normalise=True)[0]


@pytest.mark.parametrize("grp,log2n", [("g1", 16), ("g2", 13)])
def test_hot_buckets(engine, orc, grp, log2n):
    """Buckets that receive a large share of all points: equal scalars (one bucket per window), 32-bit
    scalars (rand32b of legogrothmatrix.cc:29-32: half of the points land in the carry bucket of the first
    empty window), two-valued scalars.  Exercises the warp-aggregated histogram / cursor updates and the
    multi-pass hot-bucket combine (forced tiny task lengths give thousands of partials per bucket), on the
    plain and on a precomputed key; checked with the scalar-sum identity."""
    n = 1 << log2n
    k = inputs.fr_uniform(orc, n, seed=411)
    table = engine.get_window_table(grp, 254, 0, orc.one(grp), expected_scalars=n)
    P = engine.batch_exp(254, 0, table, k)
    table.close()
    rng = np.random.default_rng(3)
    small = np.zeros((n, 4), dtype=np.uint64)
    small[:, 0] = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    two = np.where((rng.random(n) < 0.5)[:, None], inputs.fr_uniform(orc, 1, seed=412), inputs.fr_uniform(orc, 1, seed=413))
    cases = {"equal": np.tile(inputs.fr_uniform(orc, 1, seed=414), (n, 1)), "32bit": orc.fr_from_bigint(small), "two": two}
    key = engine.CommitmentKey(grp, P)
    try:
        for name, s in cases.items():
            want = _scalar_sum_expected(orc, grp, k, s)
            for c, L in ((0, 0), (12, 4), (7, 1)):
                engine.set_tuning(c, L)
                assert (key.multi_exp(s) == want).all(), (name, c, L)
            engine.set_tuning(0, 0)
            assert (engine.multi_exp(grp, P, s) == want).all(), (name, "host")
        key.precompute(11)
        engine.set_tuning_ex("use_precomputed", 2)
        for name, s in cases.items():
            assert (key.multi_exp(s) == _scalar_sum_expected(orc, grp, k, s)).all(), (name, "precomputed")
    finally:
        engine.set_tuning(0, 0)
        engine.set_tuning_ex("use_precomputed", 1)
        key.close()


@pytest.mark.parametrize("grp,log2n", [("g1", 18), ("g1", 20), ("g2", 16)])
def test_scalar_sum_identity_at_scale(engine, orc, grp, log2n):
    """sum s_i (k_i G) == (sum s_i k_i) G at BASELINE.json sizes; bases made by the GPU fixed-base path
    and spot-checked against the oracle."""
    n = 1 << log2n
    k = inputs.fr_uniform(orc, n, seed=401)
    s = inputs.fr_uniform(orc, n, seed=402)
    table = engine.get_window_table(grp, 254, 0, orc.one(grp), expected_scalars=n)
    P = engine.batch_exp(254, 0, table, k)
    table.close()
    idx = np.r_[0:64, n - 64:n, np.random.default_rng(1).integers(0, n, 256)]
    assert (P[idx] == orc.batch_exp(grp, orc.one(grp), k[idx])).all()
    want = _scalar_sum_expected(orc, grp, k, s)
    assert (engine.multi_exp(grp, P, s) == want).all()
    key = engine.CommitmentKey(grp, P)
    assert (key.multi_exp(s) == want).all()
    st = engine.last_stats()
    assert st["n"] == n and st["kernel_launches"] >= 10
    key.precompute()  # window multiples in HBM: one bucket set for all windows
    assert (key.multi_exp(s) == want).all()
    st2 = engine.last_stats()
    assert st2["num_entries"] < st["num_entries"] or log2n < 18, "precomputed key did not cut the additions"
    key.close()


@pytest.mark.parametrize("grp", ["g1", "g2"])
@pytest.mark.parametrize("levels", [1, 2])
def test_batch_affine_levels(engine, orc, golden, grp, levels):
    """Batch-affine accumulation (pair_kernels.cuh): `levels` tree levels of affine pair additions with shared
    inversions in front of the XYZZ tail must give the same group element as the plain path: every edge-case
    fixture under forced geometries (buckets of 0, 1, 2, 3, ... entries; repeated bases -> tangent case, P / -P
    pairs -> zero, zero bases), skewed scalars with hot buckets, a precomputed key, and the pipelined upload."""
    g = golden(f"msm_{grp}")
    n = 5000 if grp == "g1" else 1500
    P, k = inputs.bases(orc, grp, n, seed=801, affine=False)
    P[7] = P[6]
    P[9] = inputs.negate(orc, grp, P[8:9])[0]
    P[11] = inputs.zero_point(grp)
    s = inputs.fr_uniform(orc, n, seed=802)
    s[6:10] = s[6]  # equal scalars on the repeated / opposite bases: they meet in the same buckets
    cases = {"uniform": s, "heavy01": inputs.fr_zero_one_heavy(orc, n, seed=803), "32bit": inputs.fr_small(n, 32),
             "equal": np.tile(s[3], (n, 1))}
    want = {name: orc.msm(grp, P, v, chunks=orc.max_threads(), variant=1) for name, v in cases.items()}
    key = engine.CommitmentKey(grp, P)
    try:
        engine.set_tuning_ex("batch_affine", levels)
        for c, L in ((0, 0), (3, 32), (6, 0), (9, 64), (13, 0)):
            engine.set_tuning(c, L)
            for name, v in cases.items():
                assert (key.multi_exp(v) == want[name]).all(), (grp, levels, c, L, name)
            if c:
                for name in g["names"]:
                    B, S, R = g[f"{name}__bases"], g[f"{name}__scalars"], g[f"{name}__result"]
                    assert (engine.multi_exp(grp, B, S) == R).all(), (grp, levels, name, c, L)
        engine.set_tuning(0, 0)
        engine.set_pipeline_chunks(3)
        assert (engine.multi_exp(grp, P, cases["uniform"]) == want["uniform"]).all(), "chunks"
        engine.set_pipeline_chunks(0)
        key.precompute(9)
        engine.set_tuning_ex("use_precomputed", 2)
        for name, v in cases.items():
            assert (key.multi_exp(v) == want[name]).all(), (grp, levels, "precomputed", name)
        assert (key.multi_exp(cases["uniform"][: n // 2 + 3], offset=17) ==
                orc.msm(grp, P[17:17 + n // 2 + 3], cases["uniform"][: n // 2 + 3], chunks=orc.max_threads())).all()
    finally:
        engine.set_tuning_ex("batch_affine", 0)
        engine.set_tuning_ex("use_precomputed", 1)
        engine.set_tuning(0, 0)
        engine.set_pipeline_chunks(0)
        key.close()


def test_g2_lane_pairs(engine, orc, golden):
    """k_accumulate_g2pair (two lanes per G2 task, the halves of every Fq2 value on neighbouring lanes): same group element
    as the one-thread-per-task kernel on the edge-case fixtures under forced geometries (repeated bases -> the tangent
    branch runs on full elements inside the pair, P / -P -> zero), skewed scalars, a precomputed key and pipelined chunks."""
    g = golden("msm_g2")
    n = 2500
    P, _ = inputs.bases(orc, "g2", n, seed=821, affine=False)
    P[7] = P[6]
    P[9] = inputs.negate(orc, "g2", P[8:9])[0]
    P[11] = inputs.zero_point("g2")
    s = inputs.fr_uniform(orc, n, seed=822)
    s[6:10] = s[6]
    cases = {"uniform": s, "heavy01": inputs.fr_zero_one_heavy(orc, n, seed=823), "equal": np.tile(s[3], (n, 1))}
    want = {name: orc.msm("g2", P, v, chunks=orc.max_threads(), variant=1) for name, v in cases.items()}
    key = engine.CommitmentKey("g2", P)
    try:
        engine.set_tuning_ex("g2_lane_pairs", 1)
        for c, L in ((0, 0), (4, 32), (9, 0), (12, 64)):
            engine.set_tuning(c, L)
            for name, v in cases.items():
                assert (key.multi_exp(v) == want[name]).all(), (c, L, name)
            if c:
                for name in g["names"]:
                    B, S, R = g[f"{name}__bases"], g[f"{name}__scalars"], g[f"{name}__result"]
                    assert (engine.multi_exp("g2", B, S) == R).all(), (name, c, L)
        engine.set_tuning(0, 0)
        engine.set_pipeline_chunks(3)
        assert (engine.multi_exp("g2", P, cases["uniform"]) == want["uniform"]).all(), "chunks"
        engine.set_pipeline_chunks(0)
        key.precompute(8)
        engine.set_tuning_ex("use_precomputed", 2)
        for name, v in cases.items():
            assert (key.multi_exp(v) == want[name]).all(), ("precomputed", name)
    finally:
        engine.set_tuning_ex("g2_lane_pairs", 0)
        engine.set_tuning_ex("use_precomputed", 1)
        engine.set_tuning(0, 0)
        engine.set_pipeline_chunks(0)
        key.close()


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_alternative_kernels_under_their_knobs(engine, orc, golden, grp):
    """Kernels that stay in the tree as measured options must give the same group element as the defaults: stage 2 of the window
    reduction with one thread per partial sum (`reduce_quads` = 0; the default is the quad-cooperative adder), the mixed addition
    with paired products (`g1_paired`), the other register budgets of k_accumulate<Fq2> (`g2_blocks` = 2, 3), smaller blocks in
    stage 1 (`reduce_block`).  Edge-case fixtures under forced geometries (P == Q and P == -Q inside the adders, zero bases),
    uniform / skewed scalars on plain and precomputed keys."""
    g = golden("msm_" + grp)
    n = 3000
    P, _ = inputs.bases(orc, grp, n, seed=861, affine=False)
    P[7] = P[6]
    P[9] = inputs.negate(orc, grp, P[8:9])[0]
    P[11] = inputs.zero_point(grp)
    s = inputs.fr_uniform(orc, n, seed=862)
    s[6:10] = s[6]
    cases = {"uniform": s, "heavy01": inputs.fr_zero_one_heavy(orc, n, seed=863), "equal": np.tile(s[3], (n, 1))}
    want = {name: orc.msm(grp, P, v, chunks=orc.max_threads(), variant=1) for name, v in cases.items()}
    knobs = [("reduce_quads", 0, 1), ("reduce_block", 32, 128)]
    knobs += [("g1_paired", 1, 0)] if grp == "g1" else [("g2_blocks", 2, 1), ("g2_blocks", 3, 1)]
    key = engine.CommitmentKey(grp, P)
    pre = engine.CommitmentKey(grp, P)
    pre.precompute(8)
    try:
        for knob, value, default in knobs:
            engine.set_tuning_ex(knob, value)
            try:
                for c, L in ((0, 0), (5, 32), (11, 0)):
                    engine.set_tuning(c, L)
                    for name, v in cases.items():
                        assert (key.multi_exp(v) == want[name]).all(), (knob, value, c, L, name)
                    if c:
                        for name in g["names"]:
                            B, S, R = g[f"{name}__bases"], g[f"{name}__scalars"], g[f"{name}__result"]
                            assert (engine.multi_exp(grp, B, S) == R).all(), (knob, value, name, c, L)
                engine.set_tuning(0, 0)
                engine.set_tuning_ex("use_precomputed", 2)
                for name, v in cases.items():
                    assert (pre.multi_exp(v) == want[name]).all(), (knob, value, "precomputed", name)
            finally:
                engine.set_tuning_ex(knob, default)
                engine.set_tuning_ex("use_precomputed", 1)
                engine.set_tuning(0, 0)
    finally:
        key.close()
        pre.close()


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_reduction_meets_equal_and_opposite_sums(engine, orc, grp):
    """The adders of the window reduction must take the doubling and the cancellation branch too: with one base repeated and
    the scalars 2 and 6 (buckets 1 and 5: the same position in two neighbouring segments of four buckets) the two segment sums
    are the SAME point, so stage 2 adds a point to itself -- inside the quad-cooperative adder (default) and the serial one;
    with -P on the second half they are opposite points and the sum is the point at infinity on the way."""
    n = 6000
    P, _ = inputs.bases(orc, grp, 1, seed=871, affine=False)
    B = np.tile(P[0], (n, 1))
    s = np.tile(ints_to_mont([2], R_ORDER), (n, 1))
    s[n // 2:] = ints_to_mont([6], R_ORDER)
    Bneg = B.copy()
    Bneg[n // 2:] = inputs.negate(orc, grp, P[0:1])[0]
    s_eq = np.tile(ints_to_mont([2], R_ORDER), (n, 1))  # P and -P in the same bucket position of the same segment sums to zero
    want = orc.msm(grp, B, s, chunks=orc.max_threads(), variant=1)
    want_neg = orc.msm(grp, Bneg, s, chunks=orc.max_threads(), variant=1)
    want_zero = orc.msm(grp, Bneg, s_eq, chunks=orc.max_threads(), variant=1)
    try:
        for quads in (1, 0):
            engine.set_tuning_ex("reduce_quads", quads)
            for c, logS in ((8, 2), (8, 0), (6, 1), (10, 2)):
                engine.set_tuning(c, 0)
                engine.set_tuning_ex("reduce_log_segment", logS)
                assert (engine.multi_exp(grp, B, s) == want).all(), (quads, c, logS, "equal sums")
                assert (engine.multi_exp(grp, Bneg, s) == want_neg).all(), (quads, c, logS, "opposite sums")
                assert (engine.multi_exp(grp, Bneg, s_eq) == want_zero).all(), (quads, c, logS, "cancellation")
    finally:
        engine.set_tuning_ex("reduce_quads", 1)
        engine.set_tuning_ex("reduce_log_segment", -1)
        engine.set_tuning(0, 0)


def _device_bases(engine, orc, grp, k):
    """P_i = k_i G made by the GPU fixed-base path, spot-checked against the oracle."""
    n = len(k)
    table = engine.get_window_table(grp, 254, 0, orc.one(grp), expected_scalars=n)
    P = engine.batch_exp(254, 0, table, k)
    table.close()
    idx = np.r_[0:32, n - 32:n, np.random.default_rng(1).integers(0, n, 192)]
    assert (P[idx] == orc.batch_exp(grp, orc.one(grp), k[idx])).all()
    return P


@pytest.mark.parametrize("grp,log2n", [("g1", 22), ("g1", 24), ("g2", 18), ("g2", 20)])
def test_scalar_sum_identity_config5_sizes(engine, orc, grp, log2n):
    """BASELINE.json configs[4] sizes (G1 up to 2^24 here, G2 2^20): the host-buffer path, the plain resident
    key and the precomputed key (the engine's own window choice: c = 20 at these sizes) against the
    scalar-sum identity, whose right-hand side is computed exactly on the host (tests/inputs.py)."""
    n = 1 << log2n
    k = inputs.fr_fast_uniform(n, seed=601 + log2n)
    s = inputs.fr_fast_uniform(n, seed=602 + log2n)
    P = _device_bases(engine, orc, grp, k)
    want = inputs.scalar_sum_point(orc, grp, k, s)
    assert (engine.multi_exp(grp, P, s) == want).all(), "host-buffer path"
    key = engine.CommitmentKey(grp, P)
    try:
        del P
        assert (key.multi_exp(s) == want).all(), "plain resident key"
        key.precompute()
        assert (key.multi_exp(s) == want).all(), "precomputed key"
        assert engine.last_stats()["num_windows"] * engine.last_stats()["window_bits"] >= 254
        # a skewed vector at the same size: 32-bit scalars (rand32b, legogrothmatrix.cc:29-32)
        small = np.zeros((n, 4), dtype=np.uint64)
        small[:, 0] = np.random.default_rng(5).integers(0, 1 << 32, size=n, dtype=np.uint64)
        s32 = orc.fr_from_bigint(small)
        assert (key.multi_exp(s32) == inputs.scalar_sum_point(orc, grp, k, s32)).all(), "32-bit scalars"
    finally:
        key.close()


@pytest.mark.parametrize("grp,log2n", [("g1", 15), ("g2", 13)])
def test_precomputed_window_22(engine, orc, grp, log2n):
    """c = 22 (12 levels, 2^21 buckets) is the precomputed geometry for keys beyond 2^27 bases (13 n >= 2^31) and the
    widest one the partition sort handles; it is run here on a key small enough for the driver's suite, forced through
    b200_key_precompute_*(22)."""
    n = 1 << log2n
    k = inputs.fr_uniform(orc, n, seed=611)
    P = _device_bases(engine, orc, grp, k)
    key = engine.CommitmentKey(grp, P)
    try:
        key.precompute(22)
        engine.set_tuning_ex("use_precomputed", 2)
        for s in (inputs.fr_uniform(orc, n, seed=612), inputs.fr_zero_one_heavy(orc, n, seed=613)):
            assert (key.multi_exp(s) == inputs.scalar_sum_point(orc, grp, k, s)).all()
            st = engine.last_stats()
            assert st["window_bits"] == 22 and st["num_windows"] == 12
        m = n // 2 + 5
        s = inputs.fr_uniform(orc, m, seed=614)
        assert (key.multi_exp(s, offset=77) == inputs.scalar_sum_point(orc, grp, k[77:77 + m], s)).all()
    finally:
        engine.set_tuning_ex("use_precomputed", 1)
        key.close()


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_batched_small_msms(engine, orc, grp):
    """b200_msm_batch_*: the per-column MSMs of mtxmultiexp (LS/gadgets/subspace.cc:18-25; cplink's keygen issues 2 050
    columns of one or two terms, LS/utils/sparsemexp.h:62-90) in one call.  Columns of 0, 1, 2 and a few dozen terms,
    zero bases, zero / one scalars (sparsemexpG's special cases), repeated and opposite bases, raw Jacobian inputs;
    every column against the oracle's multi_exp."""
    rng = np.random.default_rng(9)
    sizes = [0, 1, 2, 2, 1, 0, 3, 40, 2, 1] + [int(x) for x in rng.integers(1, 3, size=60 if grp == "g1" else 20)]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    n = int(offsets[-1])
    P, _ = inputs.bases(orc, grp, n, seed=701, affine=False)
    s = inputs.fr_uniform(orc, n, seed=702)
    one = ints_to_mont([1], R_ORDER)[0]
    P[3] = inputs.zero_point(grp)
    s[4] = 0
    s[5] = one
    P[11] = P[10]                                  # column 7 holds a repeated base ...
    P[13] = inputs.negate(orc, grp, P[12:13])[0]   # ... and an opposite pair
    s[13] = s[12]
    got = engine.multi_exp_batch(grp, P, s, offsets)
    assert got.shape == (len(sizes), P.shape[1])
    for j in range(len(sizes)):
        lo, hi = int(offsets[j]), int(offsets[j + 1])
        assert (got[j] == orc.msm(grp, P[lo:hi], s[lo:hi], variant=1)).all(), (grp, j, sizes[j])
    assert engine.multi_exp_batch(grp, P[:0], s[:0], np.zeros(1, dtype=np.uint64)).shape == (0, P.shape[1])


@pytest.mark.parametrize("n", [0, 1, 65, 1026, 6000])
def test_knowledge_commitment_pair(engine, orc, n):
    """knowledge_commitment<G2,G1> MSM (SNK/knowledge_commitment/kc_multiexp.tcc:21-89: the B query
    of Groth16): both components over ONE scalar vector; sizes on both sides of the single-kernel
    threshold, 0/1-heavy scalars as witnesses have them, and the pipelined upload."""
    P2, _ = inputs.bases(orc, "g2", n, seed=501, affine=False)
    P1, _ = inputs.bases(orc, "g1", n, seed=502, affine=False)
    for s in (inputs.fr_uniform(orc, n, seed=503), inputs.fr_zero_one_heavy(orc, n)):
        o2, o1 = engine.kc_multi_exp(P2, P1, s)
        assert (o2 == orc.msm("g2", P2, s, chunks=orc.max_threads(), variant=1)).all()
        assert (o1 == orc.msm("g1", P1, s, chunks=orc.max_threads(), variant=1)).all()
    if n >= 6000:
        s = inputs.fr_uniform(orc, n, seed=504)
        try:
            engine.set_pipeline_chunks(3)
            o2, o1 = engine.kc_multi_exp(P2, P1, s)
        finally:
            engine.set_pipeline_chunks(0)
        assert (o2 == orc.msm("g2", P2, s, chunks=orc.max_threads())).all()
        assert (o1 == orc.msm("g1", P1, s, chunks=orc.max_threads())).all()
        # the G1 half must not have re-uploaded the scalars
        assert engine.last_stats()["h2d_bytes"] == n * (32 + 192 + 96)
