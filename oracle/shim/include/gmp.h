/* oracle build shim (test infrastructure): minimal declarations against libgmp.so.10 (GMP 6.3.0) ABI */
#ifndef __GMP_H__
#define __GMP_H__
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t; typedef long mp_limb_signed_t; typedef unsigned long mp_bitcnt_t;
typedef long mp_size_t; typedef long mp_exp_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1]; typedef __mpz_struct *mpz_ptr; typedef const __mpz_struct *mpz_srcptr;
typedef enum { GMP_RAND_ALG_DEFAULT = 0, GMP_RAND_ALG_LC = 0 } gmp_randalg_t;
typedef struct { mpz_t _mp_seed; gmp_randalg_t _mp_alg; union { void *_mp_lc; } _mp_algdata; } __gmp_randstate_struct;
typedef __gmp_randstate_struct gmp_randstate_t[1];
#define D(ret,name,args) ret __g##name args;
void __gmp_randinit_default(gmp_randstate_t); void __gmp_randseed_ui(gmp_randstate_t, unsigned long);
unsigned long __gmp_urandomm_ui(gmp_randstate_t, unsigned long);
void __gmp_get_memory_functions(void *(**)(size_t), void *(**)(void *, size_t, size_t), void (**)(void *, size_t));
#define gmp_randinit_default __gmp_randinit_default
#define gmp_randseed_ui __gmp_randseed_ui
#define gmp_urandomm_ui __gmp_urandomm_ui
#define Z(n) __gmpz_##n
void Z(init)(mpz_ptr); void Z(clear)(mpz_ptr); void Z(init_set)(mpz_ptr, mpz_srcptr);
void Z(set)(mpz_ptr, mpz_srcptr); void Z(set_si)(mpz_ptr, long); void Z(set_ui)(mpz_ptr, unsigned long);
void Z(set_d)(mpz_ptr, double); int Z(set_str)(mpz_ptr, const char *, int);
double Z(get_d)(mpz_srcptr); double Z(get_d_2exp)(long *, mpz_srcptr); long Z(get_si)(mpz_srcptr);
unsigned long Z(get_ui)(mpz_srcptr); char *Z(get_str)(char *, int, mpz_srcptr); size_t Z(sizeinbase)(mpz_srcptr, int);
void Z(add)(mpz_ptr, mpz_srcptr, mpz_srcptr); void Z(add_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void Z(sub)(mpz_ptr, mpz_srcptr, mpz_srcptr); void Z(sub_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void Z(mul)(mpz_ptr, mpz_srcptr, mpz_srcptr); void Z(mul_si)(mpz_ptr, mpz_srcptr, long);
void Z(mul_ui)(mpz_ptr, mpz_srcptr, unsigned long); void Z(mul_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void Z(fdiv_q_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t); void Z(mod)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void Z(addmul)(mpz_ptr, mpz_srcptr, mpz_srcptr); void Z(addmul_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void Z(submul)(mpz_ptr, mpz_srcptr, mpz_srcptr); void Z(submul_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void Z(neg)(mpz_ptr, mpz_srcptr); void Z(abs)(mpz_ptr, mpz_srcptr); void Z(swap)(mpz_ptr, mpz_ptr);
int Z(cmp)(mpz_srcptr, mpz_srcptr); int Z(cmp_si)(mpz_srcptr, long);
void Z(nextprime)(mpz_ptr, mpz_srcptr); void Z(urandomb)(mpz_ptr, gmp_randstate_t, mp_bitcnt_t);
void Z(urandomm)(mpz_ptr, gmp_randstate_t, mpz_srcptr);
#define mpz_init Z(init)
#define mpz_clear Z(clear)
#define mpz_init_set Z(init_set)
#define mpz_set Z(set)
#define mpz_set_si Z(set_si)
#define mpz_set_ui Z(set_ui)
#define mpz_set_d Z(set_d)
#define mpz_set_str Z(set_str)
#define mpz_get_d Z(get_d)
#define mpz_get_d_2exp Z(get_d_2exp)
#define mpz_get_si Z(get_si)
#define mpz_get_ui Z(get_ui)
#define mpz_get_str Z(get_str)
#define mpz_sizeinbase Z(sizeinbase)
#define mpz_add Z(add)
#define mpz_add_ui Z(add_ui)
#define mpz_sub Z(sub)
#define mpz_sub_ui Z(sub_ui)
#define mpz_mul Z(mul)
#define mpz_mul_si Z(mul_si)
#define mpz_mul_ui Z(mul_ui)
#define mpz_mul_2exp Z(mul_2exp)
#define mpz_div_2exp Z(fdiv_q_2exp)
#define mpz_mod Z(mod)
#define mpz_addmul Z(addmul)
#define mpz_addmul_ui Z(addmul_ui)
#define mpz_submul Z(submul)
#define mpz_submul_ui Z(submul_ui)
#define mpz_neg Z(neg)
#define mpz_abs Z(abs)
#define mpz_swap Z(swap)
#define mpz_cmp Z(cmp)
#define mpz_cmp_si Z(cmp_si)
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)
#define mpz_nextprime Z(nextprime)
#define mpz_urandomb Z(urandomb)
#define mpz_urandomm Z(urandomm)
#ifdef __cplusplus
}
#endif
#endif
