/* oracle build shim (test infrastructure): minimal declarations against libmpfr.so.6 (MPFR 4.2.1) ABI */
#ifndef __MPFR_H
#define __MPFR_H
#include <gmp.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef long mpfr_prec_t; typedef int mpfr_sign_t; typedef long mpfr_exp_t;
typedef struct { mpfr_prec_t _mpfr_prec; mpfr_sign_t _mpfr_sign; mpfr_exp_t _mpfr_exp; mp_limb_t *_mpfr_d; } __mpfr_struct;
typedef __mpfr_struct mpfr_t[1]; typedef __mpfr_struct *mpfr_ptr; typedef const __mpfr_struct *mpfr_srcptr;
typedef enum { MPFR_RNDN = 0, MPFR_RNDZ, MPFR_RNDU, MPFR_RNDD, MPFR_RNDA, MPFR_RNDF, MPFR_RNDNA = -1 } mpfr_rnd_t;
#define GMP_RNDN MPFR_RNDN
#define GMP_RNDZ MPFR_RNDZ
#define GMP_RNDU MPFR_RNDU
#define GMP_RNDD MPFR_RNDD
#define mp_rnd_t mpfr_rnd_t
void mpfr_init(mpfr_ptr); void mpfr_init2(mpfr_ptr, mpfr_prec_t); void mpfr_clear(mpfr_ptr);
void mpfr_set_prec(mpfr_ptr, mpfr_prec_t); void mpfr_set_default_prec(mpfr_prec_t); mpfr_prec_t mpfr_get_default_prec(void);
int mpfr_set4(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t, int);
#define mpfr_set(a, b, r) mpfr_set4(a, b, r, (b)->_mpfr_sign)
#define mpfr_init_set(x, y, r) (mpfr_init(x), mpfr_set((x), (y), (r)))
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t); int mpfr_set_ld(mpfr_ptr, long double, mpfr_rnd_t);
int mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t); int mpfr_set_z(mpfr_ptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_set_str(mpfr_ptr, const char *, int, mpfr_rnd_t); void mpfr_set_nan(mpfr_ptr); void mpfr_swap(mpfr_ptr, mpfr_ptr);
double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t); double mpfr_get_d_2exp(long *, mpfr_srcptr, mpfr_rnd_t);
long double mpfr_get_ld(mpfr_srcptr, mpfr_rnd_t); long double mpfr_get_ld_2exp(long *, mpfr_srcptr, mpfr_rnd_t);
long mpfr_get_si(mpfr_srcptr, mpfr_rnd_t); int mpfr_get_z(mpz_ptr, mpfr_srcptr, mpfr_rnd_t);
mpfr_exp_t mpfr_get_z_2exp(mpz_ptr, mpfr_srcptr);
#define mpfr_get_z_exp mpfr_get_z_2exp
mpfr_exp_t mpfr_get_exp(mpfr_srcptr); char *mpfr_get_str(char *, mpfr_exp_t *, int, size_t, mpfr_srcptr, mpfr_rnd_t);
void mpfr_free_str(char *); size_t __gmpfr_inp_str(mpfr_ptr, FILE *, int, mpfr_rnd_t);
#define mpfr_inp_str __gmpfr_inp_str
 void mpfr_free_cache(void);
int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_add_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_sub_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_mul_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_mul_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t); int mpfr_div_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_div(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_div_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_fma(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_fms(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sqrt(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_root(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_rootn_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t); int mpfr_pow_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_exp(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_log(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_hypot(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t); int mpfr_abs(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_round(mpfr_ptr, mpfr_srcptr); int mpfr_floor(mpfr_ptr, mpfr_srcptr); int mpfr_ceil(mpfr_ptr, mpfr_srcptr);
int mpfr_cmp3(mpfr_srcptr, mpfr_srcptr, int);
#define mpfr_cmp(a, b) mpfr_cmp3(a, b, 1)
int mpfr_cmp_d(mpfr_srcptr, double); int mpfr_sgn(mpfr_srcptr);
int mpfr_nan_p(mpfr_srcptr); int mpfr_number_p(mpfr_srcptr); int mpfr_zero_p(mpfr_srcptr);
#ifdef __cplusplus
}
#endif
#endif
