/* bn254_oracle.c — plain-C CPU restatement of the reference's MSM / batch_exp
 * hot path over BN254 (alt_bn128 / bn128).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker.  The product (legosnark_b200/libb200msm.so) never links
 * or calls it and has no CPU fallback.
 *
 * Parity status: PINNED against the unmodified reference sources compiled in
 * place (oracle/_ref/libffref.so, recipe in oracle/Makefile) and against the
 * committed fixtures under tests/golden/ generated from that build.
 *
 * Abbreviations: LFF = /root/reference/depends/libsnark/depends/libff/libff.
 * Representation: 4 x u64 little-endian limbs, Montgomery form, R = 2^256
 * (LFF/algebra/fields/fp.hpp:42, bigint.hpp:35).
 */
#include "bn254_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ */
/* bigint<4> helpers: LFF/algebra/fields/bigint.tcc:103-130,150-164    */
/* ------------------------------------------------------------------ */
static size_t bigint_num_bits(const uint64_t a[4])
{
    for (int i = 3; i >= 0; --i) {
        if (a[i] != 0) return (size_t)(64 * (i + 1) - __builtin_clzll(a[i]));
    }
    return 0;
}

static int bigint_test_bit(const uint64_t a[4], size_t bitno)
{
    if (bitno >= 256) return 0;
    return (int)((a[bitno >> 6] >> (bitno & 63)) & 1);
}

/* libff::log2 = ceil(log2(n)), LFF/common/utils.cpp:32-45 */
static size_t orc_log2(size_t n)
{
    size_t r = ((n & (n - 1)) == 0 ? 0 : 1);
    while (n > 1) {
        n >>= 1;
        r++;
    }
    return r;
}

/* ------------------------------------------------------------------ */
/* Fp_model<4, p>: LFF/algebra/fields/fp.tcc                            */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t l[4]; } fp_t;
typedef struct {
    uint64_t m[4]; /* modulus */
    uint64_t inv;  /* -p^{-1} mod 2^64 (alt_bn128_init.cpp:48,74) */
    fp_t one;      /* R mod p */
    fp_t r2;       /* R^2 mod p (Rsquared, alt_bn128_init.cpp:46,72) */
} fpctx_t;

/* q, r: alt_bn128_init.cpp:40,66 ; bn128_init.cpp:38,63 (identical values) */
static const fpctx_t FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}},
};
static const fpctx_t FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}},
};

static int limbs_geq(const uint64_t a[4], const uint64_t b[4])
{
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

static uint64_t limbs_sub(uint64_t o[4], const uint64_t a[4], const uint64_t b[4])
{
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        o[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

static uint64_t limbs_add(uint64_t o[4], const uint64_t a[4], const uint64_t b[4])
{
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        o[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    return carry;
}

static int fp_is_zero_raw(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }

/* mul_reduce: Montgomery product a*b*R^-1 mod p with one final conditional
 * subtraction (HAC 14.36 as in fp.tcc:161-186; CIOS ordering gives the same value). */
static void mont_mul_raw(uint64_t o[4], const uint64_t a[4], const uint64_t b[4], const fpctx_t *c)
{
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 s = (u128)a[j] * b[i] + t[j] + carry;
            t[j] = (uint64_t)s;
            carry = s >> 64;
        }
        u128 s = (u128)t[4] + carry;
        t[4] = (uint64_t)s;
        t[5] = (uint64_t)(s >> 64);
        uint64_t m = t[0] * c->inv;
        carry = ((u128)m * c->m[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            s = (u128)m * c->m[j] + t[j] + carry;
            t[j - 1] = (uint64_t)s;
            carry = s >> 64;
        }
        s = (u128)t[4] + carry;
        t[3] = (uint64_t)s;
        t[4] = t[5] + (uint64_t)(s >> 64);
    }
    if (t[4] || limbs_geq(t, c->m)) limbs_sub(t, t, c->m);
    memcpy(o, t, 32);
}

static void fp_mul(fp_t *o, const fp_t *a, const fp_t *b, const fpctx_t *c) { mont_mul_raw(o->l, a->l, b->l, c); }

/* operator+= : fp.tcc:309-420 (add, then subtract p if carry or >= p) */
static void fp_add(fp_t *o, const fp_t *a, const fp_t *b, const fpctx_t *c)
{
    uint64_t t[4];
    uint64_t carry = limbs_add(t, a->l, b->l);
    if (carry || limbs_geq(t, c->m)) limbs_sub(t, t, c->m);
    memcpy(o->l, t, 32);
}

/* operator-= : fp.tcc:422-510 (if a < b add p first) */
static void fp_sub(fp_t *o, const fp_t *a, const fp_t *b, const fpctx_t *c)
{
    uint64_t t[4];
    uint64_t borrow = limbs_sub(t, a->l, b->l);
    if (borrow) limbs_add(t, t, c->m);
    memcpy(o->l, t, 32);
}

/* operator-() : fp.tcc:551-567 (zero stays zero, else p - a) */
static void fp_neg(fp_t *o, const fp_t *a, const fpctx_t *c)
{
    if (fp_is_zero_raw(a->l)) { *o = *a; return; }
    limbs_sub(o->l, c->m, a->l);
}

/* as_bigint(): mul_reduce by the integer 1, fp.tcc:227-238 */
static void fp_as_bigint(uint64_t o[4], const uint64_t a_mont[4], const fpctx_t *c)
{
    const uint64_t one[4] = {1, 0, 0, 0};
    mont_mul_raw(o, a_mont, one, c);
}

/* Fp_model(bigint): mul_reduce(Rsquared), fp.tcc:189-194 */
static void fp_from_bigint(fp_t *o, const uint64_t a[4], const fpctx_t *c) { mont_mul_raw(o->l, a, c->r2.l, c); }

/* inverse(): the reference runs mpn_gcdext then multiplies by R^3 (fp.tcc:641-685).
 * The inverse of a field element is unique, so a^(p-2) (square-and-multiply
 * over the Montgomery representation) returns the identical limbs. */
static void fp_inv(fp_t *o, const fp_t *a, const fpctx_t *c)
{
    uint64_t e[4];
    const uint64_t two[4] = {2, 0, 0, 0};
    limbs_sub(e, c->m, two);
    fp_t r = c->one, base = *a;
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(&r, &r, &base, c);
        fp_mul(&base, &base, &base, c);
    }
    *o = r;
}

/* ---- Fq wrappers with the name shape bn254_group.inc expects ---- */
typedef fp_t fq_t;
static void fq_mul(fq_t *o, const fq_t *a, const fq_t *b) { fp_mul(o, a, b, &FQ); }
static void fq_sqr(fq_t *o, const fq_t *a) { fp_mul(o, a, a, &FQ); }
static void fq_add(fq_t *o, const fq_t *a, const fq_t *b) { fp_add(o, a, b, &FQ); }
static void fq_sub(fq_t *o, const fq_t *a, const fq_t *b) { fp_sub(o, a, b, &FQ); }
static void fq_neg(fq_t *o, const fq_t *a) { fp_neg(o, a, &FQ); }
static void fq_inv(fq_t *o, const fq_t *a) { fp_inv(o, a, &FQ); }
static int fq_is_zero(const fq_t *a) { return fp_is_zero_raw(a->l); }
static int fq_eq(const fq_t *a, const fq_t *b) { return memcmp(a->l, b->l, 32) == 0; }
static void fq_set_zero(fq_t *a) { memset(a->l, 0, 32); }
static void fq_set_one(fq_t *a) { *a = FQ.one; }

/* ------------------------------------------------------------------ */
/* Fp2_model: LFF/algebra/fields/fp2.tcc, non_residue = -1              */
/* (alt_bn128_init.cpp:95), layout c0,c1 (fp2.hpp:49)                   */
/* ------------------------------------------------------------------ */
typedef struct { fq_t c0, c1; } fq2_t;

/* Karatsuba, fp2.tcc:72-84 */
static void fq2_mul(fq2_t *o, const fq2_t *x, const fq2_t *y)
{
    fq_t aA, bB, s, t, c1;
    fq_mul(&aA, &x->c0, &y->c0);
    fq_mul(&bB, &x->c1, &y->c1);
    fq_add(&s, &x->c0, &x->c1);
    fq_add(&t, &y->c0, &y->c1);
    fq_mul(&c1, &s, &t);
    fq_sub(&c1, &c1, &aA);
    fq_sub(&c1, &c1, &bB);
    fq_sub(&o->c0, &aA, &bB); /* aA + non_residue*bB */
    o->c1 = c1;
}

/* squared_complex, fp2.tcc:111-120 */
static void fq2_sqr(fq2_t *o, const fq2_t *x)
{
    fq_t ab, s, d, c0;
    fq_mul(&ab, &x->c0, &x->c1);
    fq_add(&s, &x->c0, &x->c1);
    fq_sub(&d, &x->c0, &x->c1); /* a + non_residue*b */
    fq_mul(&c0, &s, &d);        /* (a+b)(a-b) - ab - (-ab) = (a+b)(a-b) */
    o->c0 = c0;
    fq_add(&o->c1, &ab, &ab);
}

static void fq2_add(fq2_t *o, const fq2_t *a, const fq2_t *b) { fq_add(&o->c0, &a->c0, &b->c0); fq_add(&o->c1, &a->c1, &b->c1); }
static void fq2_sub(fq2_t *o, const fq2_t *a, const fq2_t *b) { fq_sub(&o->c0, &a->c0, &b->c0); fq_sub(&o->c1, &a->c1, &b->c1); }
static void fq2_neg(fq2_t *o, const fq2_t *a) { fq_neg(&o->c0, &a->c0); fq_neg(&o->c1, &a->c1); }

/* inverse, fp2.tcc:122-136 */
static void fq2_inv(fq2_t *o, const fq2_t *x)
{
    fq_t t0, t1, t2, t3;
    fq_sqr(&t0, &x->c0);
    fq_sqr(&t1, &x->c1);
    fq_add(&t2, &t0, &t1); /* t0 - non_residue*t1 */
    fq_inv(&t3, &t2);
    fq_mul(&o->c0, &x->c0, &t3);
    fq_mul(&t0, &x->c1, &t3);
    fq_neg(&o->c1, &t0);
}
static int fq2_is_zero(const fq2_t *a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static int fq2_eq(const fq2_t *a, const fq2_t *b) { return fq_eq(&a->c0, &b->c0) && fq_eq(&a->c1, &b->c1); }
static void fq2_set_zero(fq2_t *a) { fq_set_zero(&a->c0); fq_set_zero(&a->c1); }
static void fq2_set_one(fq2_t *a) { fq_set_one(&a->c0); fq_set_zero(&a->c1); }

/* ------------------------------------------------------------------ */
/* get_exp_window_size, multiexp.tcc:509-545; tables:                   */
/* alt_bn128_init.cpp:157-201 (G1), :220-264 (G2)                       */
/* ------------------------------------------------------------------ */
static const size_t G1_WTAB[22] = {1, 5, 11, 32, 55, 162, 360, 815, 2373, 6978, 7122, 0, 57818, 0, 169679,
                                   439759, 936073, 0, 4666555, 7580404, 0, 34552892};
static const size_t G2_WTAB[22] = {1, 5, 10, 25, 59, 154, 334, 743, 2034, 4988, 8888, 26271, 39768, 106276,
                                   141703, 462423, 926872, 0, 4873049, 5706708, 0, 31673815};

static size_t orc_exp_window_size(const size_t *tab, size_t len, size_t num_scalars)
{
    size_t window = 1;
    for (long i = (long)len - 1; i >= 0; --i) {
        if (tab[i] != 0 && num_scalars >= tab[i]) {
            window = (size_t)i + 1;
            break;
        }
    }
    return window;
}

/* ------------------------------------------------------------------ */
/* group law + multiexp, instantiated for G1 (Fq) and G2 (Fq2)          */
/* ------------------------------------------------------------------ */
#define GP g1
#define FE fq
#include "bn254_group.inc"
#undef GP
#undef FE

#define GP g2
#define FE fq2
#include "bn254_group.inc"
#undef GP
#undef FE

/* ------------------------------------------------------------------ */
/* SHA-512 (FIPS 180-4) for SHA512_rng, LFF/common/rng.tcc:26-72        */
/* ------------------------------------------------------------------ */
static const uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL,
    0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL,
    0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL,
    0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL,
    0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL,
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL,
    0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL,
    0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL,
    0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL,
    0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL,
    0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL,
    0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL,
    0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

#define ROR64(x, n) (((x) >> (n)) | ((x) << (64 - (n))))

/* single-block SHA-512 of a 16-byte message (idx || iter) */
static void sha512_16(const uint8_t msg[16], uint8_t digest[64])
{
    uint64_t h[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                     0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    uint8_t blk[128];
    memset(blk, 0, 128);
    memcpy(blk, msg, 16);
    blk[16] = 0x80;
    blk[127] = 128; /* message length in bits, big-endian */
    uint64_t w[80];
    for (int i = 0; i < 16; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | blk[8 * i + j];
        w[i] = v;
    }
    for (int i = 16; i < 80; i++) {
        uint64_t s0 = ROR64(w[i - 15], 1) ^ ROR64(w[i - 15], 8) ^ (w[i - 15] >> 7);
        uint64_t s1 = ROR64(w[i - 2], 19) ^ ROR64(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i++) {
        uint64_t S1 = ROR64(e, 14) ^ ROR64(e, 18) ^ ROR64(e, 41);
        uint64_t ch = (e & f) ^ (~e & g);
        uint64_t t1 = hh + S1 + ch + K512[i] + w[i];
        uint64_t S0 = ROR64(a, 28) ^ ROR64(a, 34) ^ ROR64(a, 39);
        uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint64_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) digest[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
}

/* SHA512_rng<Fr>(idx): hash (idx, iter) little-endian, take the low 256 bits as
 * limbs, clear bits above the modulus MSB (bits 254, 255), reject if >= r,
 * then convert to Montgomery form.  rng.tcc:26-72 */
static void sha512_rng_fr(fp_t *o, uint64_t idx)
{
    uint64_t iter = 0;
    uint64_t rv[4];
    do {
        uint8_t msg[16], dg[64];
        memcpy(msg, &idx, 8);
        memcpy(msg + 8, &iter, 8);
        sha512_16(msg, dg);
        memcpy(rv, dg, 32);
        rv[3] &= 0x3fffffffffffffffULL;
        ++iter;
    } while (limbs_geq(rv, FR.m));
    fp_from_bigint(o, rv, &FR);
}

/* ------------------------------------------------------------------ */
/* exported entry points                                                */
/* ------------------------------------------------------------------ */
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks, int variant, int normalise, uint64_t *out)
{
    return g1_msm_entry(bases, scalars, n, chunks, variant, normalise, out);
}
int orc_msm_g2(const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks, int variant, int normalise, uint64_t *out)
{
    return g2_msm_entry(bases, scalars, n, chunks, variant, normalise, out);
}
int orc_batch_exp_g1(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff, size_t window, int normalise, uint64_t *out)
{
    return g1_batch_exp_entry(base, scalars, n, coeff, window, normalise, out, G1_WTAB, 22);
}
int orc_batch_exp_g2(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff, size_t window, int normalise, uint64_t *out)
{
    return g2_batch_exp_entry(base, scalars, n, coeff, window, normalise, out, G2_WTAB, 22);
}
size_t orc_exp_window_size_g1(size_t n) { return orc_exp_window_size(G1_WTAB, 22, n); }
size_t orc_exp_window_size_g2(size_t n) { return orc_exp_window_size(G2_WTAB, 22, n); }

int orc_batch_to_special_g1(uint64_t *pts, size_t n) { g1_batch_to_special((g1_t *)pts, n); return 0; }
int orc_batch_to_special_g2(uint64_t *pts, size_t n) { g2_batch_to_special((g2_t *)pts, n); return 0; }

int orc_g1_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) { return g1_op_entry(op, a, b, n, out); }
int orc_g2_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) { return g2_op_entry(op, a, b, n, out); }

int orc_scalar_mul_g1(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base, int normalise, uint64_t *out)
{
    return g1_scalar_mul_entry(base, scalars, n, stride_base, normalise, out);
}
int orc_scalar_mul_g2(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base, int normalise, uint64_t *out)
{
    return g2_scalar_mul_entry(base, scalars, n, stride_base, normalise, out);
}

static int fp_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out, const fpctx_t *c)
{
    for (size_t i = 0; i < n; i++) {
        fp_t x, y, r;
        memcpy(x.l, a + 4 * i, 32);
        if (b) memcpy(y.l, b + 4 * i, 32); else memset(y.l, 0, 32);
        switch (op) {
        case 0: fp_mul(&r, &x, &y, c); break;
        case 1: fp_mul(&r, &x, &x, c); break;
        case 2: fp_add(&r, &x, &y, c); break;
        case 3: fp_sub(&r, &x, &y, c); break;
        case 4: fp_inv(&r, &x, c); break;
        case 5: fp_neg(&r, &x, c); break;
        default: return 1;
        }
        memcpy(out + 4 * i, r.l, 32);
    }
    return 0;
}
int orc_fq_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) { return fp_op(op, a, b, n, out, &FQ); }
int orc_fr_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) { return fp_op(op, a, b, n, out, &FR); }

int orc_fq2_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    for (size_t i = 0; i < n; i++) {
        fq2_t x, y, r;
        memcpy(&x, a + 8 * i, 64);
        if (b) memcpy(&y, b + 8 * i, 64); else memset(&y, 0, 64);
        switch (op) {
        case 0: fq2_mul(&r, &x, &y); break;
        case 1: fq2_sqr(&r, &x); break;
        case 2: fq2_add(&r, &x, &y); break;
        case 3: fq2_sub(&r, &x, &y); break;
        case 4: fq2_inv(&r, &x); break;
        case 5: fq2_neg(&r, &x); break;
        default: return 1;
        }
        memcpy(out + 8 * i, &r, 64);
    }
    return 0;
}

int orc_fr_from_bigint(const uint64_t *a, size_t n, uint64_t *out)
{
    for (size_t i = 0; i < n; i++) {
        fp_t r;
        fp_from_bigint(&r, a + 4 * i, &FR);
        memcpy(out + 4 * i, r.l, 32);
    }
    return 0;
}
int orc_fr_as_bigint(const uint64_t *a, size_t n, uint64_t *out)
{
    for (size_t i = 0; i < n; i++) fp_as_bigint(out + 4 * i, a + 4 * i, &FR);
    return 0;
}
int orc_fq_from_bigint(const uint64_t *a, size_t n, uint64_t *out)
{
    for (size_t i = 0; i < n; i++) {
        fp_t r;
        fp_from_bigint(&r, a + 4 * i, &FQ);
        memcpy(out + 4 * i, r.l, 32);
    }
    return 0;
}

int orc_sha512_rng_fr(uint64_t idx0, size_t n, uint64_t *out)
{
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) {
        fp_t r;
        sha512_rng_fr(&r, idx0 + i);
        memcpy(out + 4 * i, r.l, 32);
    }
    return 0;
}


/* ------------------------------------------------------------------ */
/* Fr vector work on either side of the MSMs */
/* LS = src/, FQFFT = depends/libsnark/depends/libfqfft/libfqfft/         */
/* ------------------------------------------------------------------ */
static void fr_mul(fp_t *o, const fp_t *a, const fp_t *b) { fp_mul(o, a, b, &FR); }
static void fr_add(fp_t *o, const fp_t *a, const fp_t *b) { fp_add(o, a, b, &FR); }
static void fr_sub(fp_t *o, const fp_t *a, const fp_t *b) { fp_sub(o, a, b, &FR); }

/* CPPoly::prove, witness coefficients: LS/gadgets/poly.h:45-67.
 * w_coeffs has 2^d entries (zero-initialised, 2^d - 1 written); tmp_v ends with one value. */
int orc_fr_fold_witness(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *w_coeffs, uint64_t *eval)
{
    const size_t N = (size_t)1 << d;
    fp_t *tmp_v = (fp_t *)malloc(N * sizeof(fp_t));
    fp_t *w = w_coeffs ? (fp_t *)w_coeffs : NULL;
    const fp_t *rr = (const fp_t *)r;
    if (!tmp_v) return 1;
    memcpy(tmp_v, v, N * sizeof(fp_t));
    if (w) memset(w, 0, N * sizeof(fp_t));
    size_t start = 0;
    for (size_t i = 0; i < d; i++) {
        const size_t pBound = (size_t)1 << (d - i - 1);
        fp_t r_minus_1;
        fr_sub(&r_minus_1, &rr[i], &FR.one);
        for (size_t p = 0; p < pBound; p++) {
            const size_t p0 = p << 1, p1 = (p << 1) + 1;
            fp_t neg0, a, b;
            fp_neg(&neg0, &tmp_v[p0], &FR);
            if (w) fr_add(&w[start + p], &neg0, &tmp_v[p1]);   /* -tmp_v[p0] + tmp_v[p1]              :57 */
            fr_mul(&a, &neg0, &r_minus_1);                      /* -tmp_v[p0]*(r[i]-1) + tmp_v[p1]*r[i] :58 */
            fr_mul(&b, &tmp_v[p1], &rr[i]);
            fr_add(&tmp_v[p], &a, &b);
        }
        start += pBound;
    }
    if (eval) memcpy(eval, &tmp_v[0], sizeof(fp_t));
    free(tmp_v);
    return 0;
}

/* MultiVPolyT::evalMLE: LS/prototools/polytools.h:207-234 (table of the 2^d monomials, then a dot product) */
int orc_fr_eval_mle(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *out)
{
    const size_t N = (size_t)1 << d;
    fp_t *products = (fp_t *)malloc(N * sizeof(fp_t));
    const fp_t *vv = (const fp_t *)v, *rr = (const fp_t *)r;
    if (!products) return 1;
    products[0] = FR.one;
    size_t idx = 1;
    for (size_t i = 0; i < d; i++) {
        const size_t pBound = (size_t)1 << i;
        fp_t one_minus_r;
        fr_sub(&one_minus_r, &FR.one, &rr[i]);
        for (size_t p = 0; p < pBound; p++) {
            fr_mul(&products[p + idx], &products[p], &rr[i]);
            fr_mul(&products[p], &products[p], &one_minus_r);
        }
        idx += (size_t)1 << i;
    }
    fp_t acc, t;
    memset(&acc, 0, sizeof acc);
    for (size_t p = 0; p < N; p++) {
        fr_mul(&t, &vv[p], &products[p]);
        fr_add(&acc, &acc, &t);
    }
    memcpy(out, &acc, sizeof acc);
    free(products);
    return 0;
}

/* DPMle::pushRandomness: LS/prototools/mle.h:199-210; eqbit(false, r) = 1 - r, eqbit(true, r) = r */
int orc_fr_mle_bind(const uint64_t *table, size_t half, const uint64_t *r, uint64_t *out)
{
    const fp_t *old = (const fp_t *)table, *rr = (const fp_t *)r;
    fp_t *cur = (fp_t *)out;
    fp_t one_minus_r;
    fr_sub(&one_minus_r, &FR.one, rr);
    for (size_t p = 0; p < half; p++) {
        fp_t a, b;
        fr_mul(&a, &old[p], &one_minus_r);
        fr_mul(&b, &old[p + half], rr);
        fr_add(&cur[p], &a, &b);
    }
    return 0;
}

/* DPBeta::compute_eq_tbl: LS/prototools/mle.h:93-105, level by level as written (tmp[p] = eqbit(msb, r[j]) * dst[p >> 1]);
 * eqbit(bool, r): LS/prototools/mle.cc:13-16 */
int orc_fr_eq_table(const uint64_t *r, size_t d, uint64_t *out)
{
    const size_t N = (size_t)1 << d;
    const fp_t *rr = (const fp_t *)r;
    fp_t *dst = (fp_t *)malloc(N * sizeof(fp_t)), *tmp = (fp_t *)malloc(N * sizeof(fp_t));
    if (!dst || !tmp || d == 0) return 1;
    fr_sub(&dst[0], &FR.one, &rr[0]);
    dst[1] = rr[0];
    for (size_t j = 1; j < d; j++) {
        fp_t one_minus;
        fr_sub(&one_minus, &FR.one, &rr[j]);
        for (size_t p = 0; p < ((size_t)1 << (j + 1)); p++) {
            const int msb = p >= ((size_t)1 << j);
            fr_mul(&tmp[p], msb ? &rr[j] : &one_minus, &dst[p >> 1]);
        }
        fp_t *sw = tmp;
        tmp = dst;
        dst = sw;
    }
    memcpy(out, dst, N * sizeof(fp_t));
    free(dst);
    free(tmp);
    return 0;
}

/* DPMatrixMle::DPMatrixMle: LS/prototools/mle.h:241-259: v[r] = sum_l A[(l << d) + r] * eqTbl[l] */
int orc_fr_matrix_mle(const uint64_t *A, const uint64_t *rho, size_t d, uint64_t *v)
{
    const size_t n = (size_t)1 << d;
    const fp_t *AA = (const fp_t *)A;
    fp_t *eq = (fp_t *)malloc(n * sizeof(fp_t)), *vv = (fp_t *)v;
    if (!eq || orc_fr_eq_table(rho, d, (uint64_t *)eq)) return 1;
    memset(vv, 0, n * sizeof(fp_t));
    for (size_t rr = 0; rr < n; rr++)
        for (size_t l = 0; l < n; l++) {
            fp_t inc;
            fr_mul(&inc, &AA[(l << d) + rr], &eq[l]);
            fr_add(&vv[rr], &vv[rr], &inc);
        }
    free(eq);
    return 0;
}

/* The sum of CPSumcheck::make_new_h_poly (LS/gadgets/sumcheck.h:85-106) over two DPMle tables (getMLEPoly, mle.h:218-227:
 * eqbit_poly(0) * v0 + eqbit_poly(1) * v1 = v0 + (v1 - v0) x), polynomial products as PolyT::mul (polytools.h:54-64), with an
 * optional per-p scalar w[p] (the beta suffix; NULL = DPBetaDummy).  out = 3 coefficients. */
int orc_fr_sumcheck_round(const uint64_t *a, const uint64_t *b, const uint64_t *w, size_t half, uint64_t *out)
{
    const fp_t *aa = (const fp_t *)a, *bb = (const fp_t *)b, *ww = (const fp_t *)w;
    fp_t c[3], minus_one;
    memset(c, 0, sizeof c);
    fr_sub(&minus_one, &c[0], &FR.one);
    for (size_t p = 0; p < half; p++) {
        fp_t pa[2], pb[2], t, prod[3];
        /* mle poly: (1, -1) * v0 + (0, 1) * v1 */
        pa[0] = aa[p];
        fr_mul(&t, &minus_one, &aa[p]);
        fr_add(&pa[1], &t, &aa[p + half]);
        pb[0] = bb[p];
        fr_mul(&t, &minus_one, &bb[p]);
        fr_add(&pb[1], &t, &bb[p + half]);
        if (ww) {
            fr_mul(&pa[0], &pa[0], &ww[p]);
            fr_mul(&pa[1], &pa[1], &ww[p]);
        }
        memset(prod, 0, sizeof prod);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) {
                fr_mul(&t, &pa[i], &pb[j]);
                fr_add(&prod[i + j], &prod[i + j], &t);
            }
        for (int k = 0; k < 3; k++) fr_add(&c[k], &c[k], &prod[k]);
    }
    memcpy(out, c, sizeof c);
    return 0;
}

/* the round loop of CPSumcheck::prove (LS/gadgets/sumcheck.cc:56-70) without a beta factor: h[i], then pushRandomness(r[i]) */
int orc_fr_sumcheck_rounds(const uint64_t *a, const uint64_t *b, const uint64_t *r, size_t d, uint64_t *h)
{
    const size_t N = (size_t)1 << d;
    uint64_t *ca = (uint64_t *)malloc(N * 32), *cb = (uint64_t *)malloc(N * 32), *na = (uint64_t *)malloc(N * 32), *nb = (uint64_t *)malloc(N * 32);
    if (!ca || !cb || !na || !nb) return 1;
    memcpy(ca, a, N * 32);
    memcpy(cb, b, N * 32);
    for (size_t i = 0; i < d; i++) {
        const size_t half = (size_t)1 << (d - i - 1);
        orc_fr_sumcheck_round(ca, cb, NULL, half, h + 12 * i);
        if (i + 1 < d) {
            orc_fr_mle_bind(ca, half, r + 4 * i, na);
            orc_fr_mle_bind(cb, half, r + 4 * i, nb);
            uint64_t *t = ca; ca = na; na = t;
            t = cb; cb = nb; nb = t;
        }
    }
    free(ca); free(cb); free(na); free(nb);
    return 0;
}

/* Fr::root_of_unity, Fr::s = 28 (alt_bn128_init.cpp:57-60); get_root_of_unity: field_utils.tcc:38-51 */
static void fr_root_of_unity(fp_t *omega, size_t logn)
{
    /* 19103219067921713944291392827692070036145651957329286315305642004821462161904 */
    static const uint64_t ROOT[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};
    fp_from_bigint(omega, ROOT, &FR);
    for (size_t i = 28; i > logn; --i) fr_mul(omega, omega, omega);
}

static void fr_pow_u64(fp_t *o, const fp_t *base, uint64_t e)
{
    fp_t r = FR.one, b = *base;
    while (e) {
        if (e & 1) fr_mul(&r, &r, &b);
        fr_mul(&b, &b, &b);
        e >>= 1;
    }
    *o = r;
}

/* _basic_serial_radix2_FFT: FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.tcc:42-75 */
static void fr_serial_radix2_fft(fp_t *a, size_t n, const fp_t *omega)
{
    const size_t logn = orc_log2(n);
    for (size_t k = 0; k < n; ++k) {
        size_t rk = 0;
        for (size_t b = 0; b < logn; b++) rk |= ((k >> b) & 1) << (logn - 1 - b);  /* libff::bitreverse */
        if (k < rk) {
            const fp_t t = a[k];
            a[k] = a[rk];
            a[rk] = t;
        }
    }
    size_t m = 1;
    for (size_t s = 1; s <= logn; ++s) {
        fp_t w_m;
        fr_pow_u64(&w_m, omega, n / (2 * m));
        for (size_t k = 0; k < n; k += 2 * m) {
            fp_t w = FR.one;
            for (size_t j = 0; j < m; ++j) {
                fp_t t;
                fr_mul(&t, &w, &a[k + j + m]);
                fr_sub(&a[k + j + m], &a[k + j], &t);
                fr_add(&a[k + j], &a[k + j], &t);
                fr_mul(&w, &w, &w_m);
            }
        }
        m *= 2;
    }
}

/* _multiply_by_coset: basic_radix2_domain_aux.tcc:163-171 */
static void fr_multiply_by_coset(fp_t *a, size_t n, const fp_t *g)
{
    fp_t u = *g;
    for (size_t i = 1; i < n; ++i) {
        fr_mul(&a[i], &a[i], &u);
        fr_mul(&u, &u, g);
    }
}

/* basic_radix2_domain<Fr>::FFT / iFFT / cosetFFT / icosetFFT: basic_radix2_domain.tcc:41-78.
 * mode 0 FFT, 1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g), 4 unscaled inverse; a has 2^log_n entries, in place. */
int orc_fr_fft(uint64_t *a_, size_t log_n, int mode, const uint64_t *g_)
{
    fp_t *a = (fp_t *)a_;
    const size_t n = (size_t)1 << log_n;
    fp_t omega, g;
    if (log_n < 1 || log_n > 28 || mode < 0 || mode > 4 || ((mode == 2 || mode == 3) && !g_)) return 1;
    fr_root_of_unity(&omega, log_n);
    if (g_) memcpy(&g, g_, sizeof g);
    if (mode == 2) fr_multiply_by_coset(a, n, &g);
    if (mode == 0 || mode == 2) {
        fr_serial_radix2_fft(a, n, &omega);
        return 0;
    }
    fp_t omega_inv, sconst, nn;
    fp_inv(&omega_inv, &omega, &FR);
    fr_serial_radix2_fft(a, n, &omega_inv);
    if (mode == 4) return 0; /* _basic_radix2_FFT(a, omega.inverse()) alone: what extended / step domains call */
    const uint64_t nb[4] = {(uint64_t)n, 0, 0, 0};
    fp_from_bigint(&nn, nb, &FR);
    fp_inv(&sconst, &nn, &FR);
    for (size_t i = 0; i < n; ++i) fr_mul(&a[i], &a[i], &sconst);
    if (mode == 3) {
        fp_t ginv;
        fp_inv(&ginv, &g, &FR);
        fr_multiply_by_coset(a, n, &ginv);
    }
    return 0;
}


/* step_radix2_domain<Fr>::FFT / iFFT / cosetFFT / icosetFFT: FQFFT/evaluation_domain/domains/step_radix2_domain.tcc:38-152
 * (FQFFT = depends/libsnark/depends/libfqfft/libfqfft), domains of big + small = 2^log_big + 2^log_small points.
 * mode 0 FFT, 1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g); in place.  The radix-2 transforms inside are orc_fr_fft modes 0 / 4
 * (_basic_radix2_FFT with omega^2 resp. the small root, and their inverses). */
int orc_fr_step_fft(uint64_t *a_, size_t log_big, size_t log_small, int mode, const uint64_t *g_)
{
    fp_t *a = (fp_t *)a_;
    const size_t big = (size_t)1 << log_big, small = (size_t)1 << log_small, m = big + small, compr = big / small;
    if (log_small >= log_big || log_big + 1 > 28 || mode < 0 || mode > 3 || (mode >= 2 && !g_)) return 1;
    fp_t omega, g;
    fr_root_of_unity(&omega, log_big + 1); /* :31: get_root_of_unity(1 << log2(m)), log2 = ceil */
    if (g_) memcpy(&g, g_, sizeof g);
    if (mode == 0 || mode == 2) {
        if (mode == 2) fr_multiply_by_coset(a, m, &g); /* :142-146 */
        fp_t *c = (fp_t *)malloc(big * sizeof(fp_t)), *d = (fp_t *)malloc(big * sizeof(fp_t)), *e = (fp_t *)calloc(small, sizeof(fp_t));
        if (!c || !d || !e) return 1;
        fp_t omega_i = FR.one, t;
        for (size_t i = 0; i < big; ++i) { /* :43-50 */
            if (i < small) {
                fr_add(&c[i], &a[i], &a[i + big]);
                fr_sub(&t, &a[i], &a[i + big]);
            } else {
                c[i] = a[i];
                t = a[i];
            }
            fr_mul(&d[i], &omega_i, &t);
            fr_mul(&omega_i, &omega_i, &omega);
        }
        for (size_t i = 0; i < small; ++i) /* :52-60 */
            for (size_t j = 0; j < compr; ++j) fr_add(&e[i], &e[i], &d[i + j * small]);
        orc_fr_fft((uint64_t *)c, log_big, 0, NULL);                  /* _basic_radix2_FFT(c, omega.squared()) */
        if (log_small >= 1) orc_fr_fft((uint64_t *)e, log_small, 0, NULL); /* size 1: the identity */
        memcpy(a, c, big * sizeof(fp_t));
        memcpy(a + big, e, small * sizeof(fp_t));
        free(c); free(d); free(e);
        return 0;
    }
    /* iFFT :73-139 */
    fp_t *U0 = (fp_t *)malloc(big * sizeof(fp_t)), *U1 = (fp_t *)malloc(small * sizeof(fp_t)), *tmp = (fp_t *)malloc(big * sizeof(fp_t));
    if (!U0 || !U1 || !tmp) return 1;
    memcpy(U0, a, big * sizeof(fp_t));
    memcpy(U1, a + big, small * sizeof(fp_t));
    orc_fr_fft((uint64_t *)U0, log_big, 1, NULL); /* unscaled inverse, then * big^-1 (:80-87) = the scaled iFFT */
    if (log_small >= 1) orc_fr_fft((uint64_t *)U1, log_small, 1, NULL);
    fp_t omega_i = FR.one;
    for (size_t i = 0; i < big; ++i) { /* :95-101 */
        fr_mul(&tmp[i], &U0[i], &omega_i);
        fr_mul(&omega_i, &omega_i, &omega);
    }
    for (size_t i = small; i < big; ++i) a[i] = U0[i]; /* :104-107 */
    for (size_t i = 0; i < small; ++i)                   /* :110-116 */
        for (size_t j = 1; j < compr; ++j) fr_sub(&U1[i], &U1[i], &tmp[i + j * small]);
    fp_t omega_inv, omega_inv_i = FR.one, two, over_two, t;
    fp_inv(&omega_inv, &omega, &FR);
    for (size_t i = 0; i < small; ++i) { /* :118-124 */
        fr_mul(&U1[i], &U1[i], &omega_inv_i);
        fr_mul(&omega_inv_i, &omega_inv_i, &omega_inv);
    }
    fr_add(&two, &FR.one, &FR.one);
    fp_inv(&over_two, &two, &FR);
    for (size_t i = 0; i < small; ++i) { /* :127-138 */
        fr_add(&t, &U0[i], &U1[i]);
        fr_mul(&a[i], &t, &over_two);
        fr_sub(&t, &U0[i], &U1[i]);
        fr_mul(&a[big + i], &t, &over_two);
    }
    free(U0); free(U1); free(tmp);
    if (mode == 3) { /* :148-152 */
        fp_t ginv;
        fp_inv(&ginv, &g, &FR);
        fr_multiply_by_coset(a, m, &ginv);
    }
    return 0;
}

/* _basic_radix2_evaluate_all_lagrange_polynomials: FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.tcc:183-236.
 * u has m = 2^log_m entries; returns through u the values L_{i,S}(t) on S = {omega^i}. */
static void fr_basic_lagrange(fp_t *u, size_t log_m, const fp_t *t)
{
    const size_t m = (size_t)1 << log_m;
    if (m == 1) { /* :185-188 */
        u[0] = FR.one;
        return;
    }
    fp_t omega, tm;
    fr_root_of_unity(&omega, log_m);
    memset(u, 0, m * sizeof(fp_t));
    fr_pow_u64(&tm, t, m);
    if (memcmp(&tm, &FR.one, sizeof tm) == 0) { /* :201-214: t is a point of S */
        fp_t omega_i = FR.one;
        for (size_t i = 0; i < m; ++i) {
            if (memcmp(&omega_i, t, sizeof omega_i) == 0) {
                u[i] = FR.one;
                return;
            }
            fr_mul(&omega_i, &omega_i, &omega);
        }
    }
    fp_t Z, l, r = FR.one, mm, d; /* :225-233 */
    const uint64_t mb[4] = {(uint64_t)m, 0, 0, 0};
    fr_sub(&Z, &tm, &FR.one);
    fp_from_bigint(&mm, mb, &FR);
    fp_inv(&mm, &mm, &FR);
    fr_mul(&l, &Z, &mm);
    for (size_t i = 0; i < m; ++i) {
        fr_sub(&d, t, &r);
        fp_inv(&d, &d, &FR);
        fr_mul(&u[i], &l, &d);
        fr_mul(&l, &l, &omega);
        fr_mul(&r, &r, &omega);
    }
}

/* evaluate_all_lagrange_polynomials(t) of basic_radix2_domain (basic_radix2_domain.tcc:80-84; log_small == (size_t)-1) and of
 * step_radix2_domain (step_radix2_domain.tcc:161-186; 2^log_big + 2^log_small points): what libsnark's generators call through
 * r1cs_to_qap_instance_map_with_evaluation (r1cs_to_qap.tcc:127-190). */
int orc_fr_lagrange(uint64_t *out_, size_t log_big, size_t log_small, const uint64_t *t_)
{
    fp_t *out = (fp_t *)out_, t;
    memcpy(&t, t_, sizeof t);
    if (log_big > 27) return 1;
    if (log_small == (size_t)-1) {
        fr_basic_lagrange(out, log_big, &t);
        return 0;
    }
    if (log_small >= log_big) return 1;
    const size_t big = (size_t)1 << log_big, small = (size_t)1 << log_small;
    fp_t omega, big_omega, omega_inv, t_small;
    fr_root_of_unity(&omega, log_big + 1); /* step_radix2_domain.tcc:31-33: omega, big_omega = omega^2 */
    fr_mul(&big_omega, &omega, &omega);
    fp_t *inner_big = (fp_t *)malloc(big * sizeof(fp_t)), *inner_small = (fp_t *)malloc(small * sizeof(fp_t));
    if (!inner_big || !inner_small) return 1;
    fr_basic_lagrange(inner_big, log_big, &t); /* :163 */
    fp_inv(&omega_inv, &omega, &FR);
    fr_mul(&t_small, &t, &omega_inv);
    fr_basic_lagrange(inner_small, log_small, &t_small); /* :164 */
    fp_t L0, w, rho, elt = FR.one, d, x; /* :168-176 */
    fr_pow_u64(&w, &omega, small);
    fr_pow_u64(&L0, &t, small);
    fr_sub(&L0, &L0, &w);
    fr_pow_u64(&rho, &big_omega, small);
    for (size_t i = 0; i < big; ++i) {
        fr_sub(&d, &elt, &w);
        fp_inv(&d, &d, &FR);
        fr_mul(&x, &inner_big[i], &L0);
        fr_mul(&out[i], &x, &d);
        fr_mul(&elt, &elt, &rho);
    }
    fp_t L1, den; /* :178-183 */
    fr_pow_u64(&L1, &t, big);
    fr_sub(&L1, &L1, &FR.one);
    fr_pow_u64(&den, &omega, big);
    fr_sub(&den, &den, &FR.one);
    fp_inv(&den, &den, &FR);
    fr_mul(&L1, &L1, &den);
    for (size_t i = 0; i < small; ++i) fr_mul(&out[big + i], &L1, &inner_small[i]);
    free(inner_big);
    free(inner_small);
    return 0;
}

/* ------------------------------------------------------------------ */
/* wire format: point compression               */
/* operator<< / operator>> of alt_bn128_G1 (alt_bn128_g1.cpp:404-459),  */
/* alt_bn128_G2 (alt_bn128_g2.cpp:414-475), bn128_G1 (bn128_g1.cpp:     */
/* 344-463), bn128_G2 (bn128_g2.cpp:374-470), compression on.           */
/* flavour 0: alt_bn128; 1: alt_bn128 -DMONTGOMERY_OUTPUT; 2: bn128.    */
/* ------------------------------------------------------------------ */
static void fq_pow_limbs(fq_t *o, const fq_t *a, const uint64_t *e, size_t nlimbs)
{
    fq_t r;
    fq_set_one(&r);
    for (size_t i = nlimbs * 64; i-- > 0;) {
        fq_sqr(&r, &r);
        if ((e[i / 64] >> (i % 64)) & 1) fq_mul(&r, &r, a);
    }
    *o = r;
}
static void fq2_pow_limbs(fq2_t *o, const fq2_t *a, const uint64_t *e, size_t nlimbs)
{
    fq2_t r;
    fq2_set_one(&r);
    for (size_t i = nlimbs * 64; i-- > 0;) {
        fq2_sqr(&r, &r);
        if ((e[i / 64] >> (i % 64)) & 1) fq2_mul(&r, &r, a);
    }
    *o = r;
}
/* Fp_model::sqrt (fp.tcc Tonelli-Shanks) for Fq: s = 1, t_minus_1_over_2 from alt_bn128_init.cpp:82-84;
 * b = a^t is 1 for every square, so the loop body never runs and x = a * a^((t-1)/2) is returned */
static void fq_sqrt(fq_t *o, const fq_t *a)
{
    /* 5472060717959818805561601436314318772174077789324455915672259473661306552145 */
    static const uint64_t T12[4] = {0x4f082305b61f3f51ULL, 0x65e05aa45a1c72a3ULL, 0x6e14116da0605617ULL, 0x0c19139cb84c680aULL};
    fq_t w;
    fq_pow_limbs(&w, a, T12, 4);
    fq_mul(o, a, &w);
}
/* Fp2_model::sqrt, fp2.tcc:146-200; s = 4, t_minus_1_over_2 and nqr_to_t from alt_bn128_init.cpp:92-98 */
static void fq2_sqrt(fq2_t *o, const fq2_t *a)
{
    static const uint64_t T12[8] = {0x09daa2c5113aeb4dULL, 0xe5301039684f5608ULL, 0x425280c4e36cb656ULL, 0x682344f4abd09216ULL,
                                    0x31376fd2e1a6359cULL, 0xe5805c2a88b1bab0ULL, 0xe2ccd37be01a4690ULL, 0x00492e25c3b1e5fcULL};
    /* nqr_to_t = (5033503716262624267312492558379982687175200734934877598599011485707452665730,
     *             314498342015008975724433667930697407966947188435857772134235984660852259084) */
    static const uint64_t Z0[4] = {0x47cfbbedda71cf82ULL, 0x5398a41a4e1dc5d3ULL, 0x0dd3ecd4f3051527ULL, 0x0b20dcb5704e326aULL};
    static const uint64_t Z1[4] = {0xab0f3a6ca462390cULL, 0xf05cfc50e9715370ULL, 0x2252522c29527d19ULL, 0x00b1ffefd8885bf2ULL};
    fq2_t one, z, w, x, b;
    fq2_set_one(&one);
    fp_from_bigint(&z.c0, Z0, &FQ);
    fp_from_bigint(&z.c1, Z1, &FQ);
    size_t v = 4;
    fq2_pow_limbs(&w, a, T12, 8);
    fq2_mul(&x, a, &w);
    fq2_mul(&b, &x, &w);
    for (int guard = 0; guard < 8 && !fq2_eq(&b, &one); guard++) {
        size_t m = 0;
        fq2_t b2m = b;
        while (!fq2_eq(&b2m, &one) && m < 8) {
            fq2_sqr(&b2m, &b2m);
            m += 1;
        }
        int j = (int)v - (int)m - 1;
        w = z;
        while (j > 0) {
            fq2_sqr(&w, &w);
            --j;
        }
        fq2_sqr(&z, &w);
        fq2_mul(&b, &b, &z);
        fq2_mul(&x, &x, &w);
        v = m;
    }
    *o = x;
}

static void fq_to_wire(uint64_t o[4], const fq_t *x, int fl)
{
    if (fl == 0) fp_as_bigint(o, x->l, &FQ);
    else memcpy(o, x->l, 32);
}
static void fq_from_wire(fq_t *o, const uint64_t x[4], int fl)
{
    if (fl == 0) fp_from_bigint(o, x, &FQ);
    else memcpy(o->l, x, 32);
}
static unsigned fq_parity(const fq_t *y, int fl)
{
    uint64_t t[4];
    if (fl == 2) return (unsigned)(y->l[0] & 1);
    fp_as_bigint(t, y->l, &FQ);
    return (unsigned)(t[0] & 1);
}

int orc_compress_g1(const uint64_t *pts, size_t n, int fl, uint64_t *x_out, uint8_t *flags)
{
    for (size_t i = 0; i < n; i++) {
        g1_t p;
        memcpy(&p, pts + 12 * i, sizeof p);
        if (g1_is_zero(&p)) { /* to_affine_coordinates() turns a zero into (0, 1, 0) on both curves (alt_bn128_g1.cpp:60-67, bn128_g1.cpp:106-112) */
            fq_t one = FQ.one, zero;
            fq_set_zero(&zero);
            fq_to_wire(x_out + 4 * i, &zero, fl);
            flags[i] = (uint8_t)(2u | fq_parity(&one, fl));
            continue;
        }
        g1_to_affine(&p);
        fq_to_wire(x_out + 4 * i, &p.X, fl);
        flags[i] = (uint8_t)fq_parity(&p.Y, fl);
    }
    return 0;
}
int orc_compress_g2(const uint64_t *pts, size_t n, int fl, uint64_t *x_out, uint8_t *flags)
{
    for (size_t i = 0; i < n; i++) {
        g2_t p;
        memcpy(&p, pts + 24 * i, sizeof p);
        if (g2_is_zero(&p)) {
            fq_t one = FQ.one, zero;
            fq_set_zero(&zero);
            fq_to_wire(x_out + 8 * i, &zero, fl);
            fq_to_wire(x_out + 8 * i + 4, &zero, fl);
            flags[i] = (uint8_t)(2u | fq_parity(&one, fl));
            continue;
        }
        g2_to_affine(&p);
        fq_to_wire(x_out + 8 * i, &p.X.c0, fl);
        fq_to_wire(x_out + 8 * i + 4, &p.X.c1, fl);
        flags[i] = (uint8_t)fq_parity(&p.Y.c0, fl);
    }
    return 0;
}
int orc_decompress_g1(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts_out)
{
    const uint64_t three[4] = {3, 0, 0, 0};
    fq_t b;
    fp_from_bigint(&b, three, &FQ); /* alt_bn128_coeff_b, alt_bn128_init.cpp:134 */
    for (size_t i = 0; i < n; i++) {
        g1_t g;
        if (flags[i] & 2) {
            g1_set_zero(&g);
        } else {
            fq_t x2, y2;
            fq_from_wire(&g.X, x + 4 * i, fl);
            fq_sqr(&x2, &g.X);
            fq_mul(&y2, &x2, &g.X);
            fq_add(&y2, &y2, &b);
            fq_sqrt(&g.Y, &y2);
            if (fq_parity(&g.Y, fl) != (unsigned)(flags[i] & 1)) fq_neg(&g.Y, &g.Y);
            fq_set_one(&g.Z);
        }
        memcpy(pts_out + 12 * i, &g, sizeof g);
    }
    return 0;
}
int orc_decompress_g2(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts_out)
{
    /* alt_bn128_twist_coeff_b = 3 / (9 + u), alt_bn128_init.cpp:135-136 */
    static const uint64_t B0[4] = {0x3267e6dc24a138e5ULL, 0xb5b4c5e559dbefa3ULL, 0x81be18991be06ac3ULL, 0x2b149d40ceb8aaaeULL};
    static const uint64_t B1[4] = {0xe4a2bd0685c315d2ULL, 0xa74fa084e52d1852ULL, 0xcd2cafadeed8fdf4ULL, 0x009713b03af0fed4ULL};
    fq2_t b;
    fp_from_bigint(&b.c0, B0, &FQ);
    fp_from_bigint(&b.c1, B1, &FQ);
    for (size_t i = 0; i < n; i++) {
        g2_t g;
        if (flags[i] & 2) {
            g2_set_zero(&g);
        } else {
            fq2_t x2, y2;
            fq_from_wire(&g.X.c0, x + 8 * i, fl);
            fq_from_wire(&g.X.c1, x + 8 * i + 4, fl);
            fq2_sqr(&x2, &g.X);
            fq2_mul(&y2, &x2, &g.X);
            fq2_add(&y2, &y2, &b);
            fq2_sqrt(&g.Y, &y2);
            if (fq_parity(&g.Y.c0, fl) != (unsigned)(flags[i] & 1)) fq2_neg(&g.Y, &g.Y);
            fq2_set_one(&g.Z);
        }
        memcpy(pts_out + 24 * i, &g, sizeof g);
    }
    return 0;
}

/* G1_one = (1,2,1): alt_bn128_init.cpp:148-150 ; G2_one: :209-213 */
int orc_g1_one(uint64_t *out)
{
    g1_t g;
    const uint64_t one[4] = {1, 0, 0, 0}, two[4] = {2, 0, 0, 0};
    fp_from_bigint(&g.X, one, &FQ);
    fp_from_bigint(&g.Y, two, &FQ);
    fq_set_one(&g.Z);
    memcpy(out, &g, sizeof g);
    return 0;
}

int orc_g2_one(uint64_t *out)
{
    /* decimal constants of alt_bn128_init.cpp:209-212 as little-endian limbs */
    static const uint64_t XC0[4] = {0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL};
    static const uint64_t XC1[4] = {0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL};
    static const uint64_t YC0[4] = {0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL};
    static const uint64_t YC1[4] = {0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL};
    g2_t g;
    fp_from_bigint(&g.X.c0, XC0, &FQ);
    fp_from_bigint(&g.X.c1, XC1, &FQ);
    fp_from_bigint(&g.Y.c0, YC0, &FQ);
    fp_from_bigint(&g.Y.c1, YC1, &FQ);
    fq2_set_one(&g.Z);
    memcpy(out, &g, sizeof g);
    return 0;
}
