/* Declaration-only GMP ABI header for building the reference libff as the
 * parity oracle (oracle/_ref).  TEST INFRASTRUCTURE ONLY.
 *
 * The image ships the GMP 6.3.0 runtime (libgmp.so.10) but no development
 * headers.  This file declares exactly the mpn_/mpz_ entry points libff and
 * libfqfft call and links against the system runtime with -l:libgmp.so.10.
 * Functions that are macros/inlines in the real gmp.h (mpn_add_1, mpn_sub_1,
 * mpn_sub, mpz_sgn) are provided inline on top of exported symbols.
 */
#ifndef B200_ORACLE_GMP_ABI_H
#define B200_ORACLE_GMP_ABI_H

#include <stddef.h>
#include <stdio.h>

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;

#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64
#define GMP_NAIL_BITS 0
#define __GNU_MP_VERSION 6
#define __GNU_MP_VERSION_MINOR 3
#define __GNU_MP_VERSION_PATCHLEVEL 0

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#ifdef __cplusplus
extern "C" {
#endif

#define mpn_mul_n __gmpn_mul_n
#define mpn_mul __gmpn_mul
#define mpn_addmul_1 __gmpn_addmul_1
#define mpn_add_n __gmpn_add_n
#define mpn_sub_n __gmpn_sub_n
#define mpn_cmp __gmpn_cmp
#define mpn_copyi __gmpn_copyi
#define mpn_zero __gmpn_zero
#define mpn_gcdext __gmpn_gcdext
#define mpn_tdiv_qr __gmpn_tdiv_qr
#define mpn_set_str __gmpn_set_str
#define mpn_rshift __gmpn_rshift
#define mpn_lshift __gmpn_lshift
#define mpz_init __gmpz_init
#define mpz_init_set __gmpz_init_set
#define mpz_clear __gmpz_clear
#define mpz_set_ui __gmpz_set_ui
#define mpz_get_ui __gmpz_get_ui
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_add_ui __gmpz_add_ui
#define gmp_printf __gmp_printf

void mpn_mul_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_mul(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t mpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
int mpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
void mpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
void mpn_zero(mp_ptr, mp_size_t);
mp_size_t mpn_gcdext(mp_ptr, mp_ptr, mp_size_t *, mp_ptr, mp_size_t, mp_ptr, mp_size_t);
void mpn_tdiv_qr(mp_ptr, mp_ptr, mp_size_t, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_size_t mpn_set_str(mp_ptr, const unsigned char *, size_t, int);
mp_limb_t mpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t mpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);

void mpz_init(mpz_ptr);
void mpz_init_set(mpz_ptr, mpz_srcptr);
void mpz_clear(mpz_ptr);
void mpz_set_ui(mpz_ptr, unsigned long);
unsigned long mpz_get_ui(mpz_srcptr);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
int gmp_printf(const char *, ...);

#ifdef __cplusplus
}
#endif

/* header-inline in real GMP (not exported by libgmp.so) */
static inline mp_limb_t mpn_add_1(mp_ptr rp, mp_srcptr up, mp_size_t n, mp_limb_t v)
{
    mp_size_t i;
    for (i = 0; i < n; i++) {
        mp_limb_t s = up[i] + v;
        v = (s < v);
        rp[i] = s;
    }
    return v;
}
static inline mp_limb_t mpn_sub_1(mp_ptr rp, mp_srcptr up, mp_size_t n, mp_limb_t v)
{
    mp_size_t i;
    for (i = 0; i < n; i++) {
        mp_limb_t u = up[i];
        rp[i] = u - v;
        v = (u < v);
    }
    return v;
}
static inline mp_limb_t mpn_sub(mp_ptr rp, mp_srcptr up, mp_size_t un, mp_srcptr vp, mp_size_t vn)
{
    mp_limb_t b = mpn_sub_n(rp, up, vp, vn);
    if (un > vn) b = mpn_sub_1(rp + vn, up + vn, un - vn, b);
    return b;
}
static inline int mpz_sgn(mpz_srcptr z)
{
    return z->_mp_size < 0 ? -1 : (z->_mp_size > 0);
}

#endif
