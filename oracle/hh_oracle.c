/* hh_oracle.c — CPU ORACLE for the Householder path, TEST INFRASTRUCTURE ONLY (see gso_oracle.c header).
 *
 * Restates MatHouseholder<Z_NR<long>, FP_NR<double>> (fplll/householder.{h,cpp}) with HOUSEHOLDER_ROW_EXPO:
 * refresh_R_bf (householder.cpp:186-245), refresh_R (:247-261), update_R (:151-184), update_R_last (:27-146, the
 * default !HOUSEHOLDER_PRECOMPUTE_INVERSE branch), size_reduce (:403-451) with row_addmul_we (:522-559),
 * swap (:372-398), recover_R (householder.h:597-608).  Dot products and axpys keep the reference's element order
 * (numvect.h:385-395 ascending; NumVect::addmul numvect.h:300-305 descending), unfused multiply-add.
 * R_history is the reference's full per-row trace R_history[i][j][k] (householder.h:103-109): a vector that moves
 * down several positions by consecutive swaps needs the snapshot taken at j = (its new index - 1), so the
 * "diagonal + last snapshot" reduction suggested in SURVEY §7 step 7 is NOT sufficient in general.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
  int d, n, enable_row_expo, n_known_rows, n_known_cols, updated_R;
  int64_t *b;                 /* d*n */
  double *bf, *R, *V;         /* d*n */
  double *sigma;              /* d */
  double *hist;               /* d*n*n : hist[(i*n+j)*n+k] = R_history[i][j][k], k >= j */
  int64_t *row_expo;          /* d */
  int *init_row_size;         /* d */
  double *norm_square_b;      /* d */
  int64_t *expo_norm_square_b; /* d */
} ohh_t;

#define HB(i, j) m->b[(size_t)(i) * m->n + (j)]
#define HBF(i, j) m->bf[(size_t)(i) * m->n + (j)]
#define HR(i, j) m->R[(size_t)(i) * m->n + (j)]
#define HV(i, j) m->V[(size_t)(i) * m->n + (j)]

static int hmax(int a, int b) { return a > b ? a : b; }

ohh_t *ohh_create(int d, int n, const int64_t *b, int flags)
{
  ohh_t *m           = (ohh_t *)calloc(1, sizeof(ohh_t));
  m->d = d, m->n = n, m->enable_row_expo = (flags & 1) ? 1 : 0; /* HOUSEHOLDER_ROW_EXPO = 1, householder.h:26-32 */
  m->b                  = (int64_t *)malloc(sizeof(int64_t) * d * n);
  memcpy(m->b, b, sizeof(int64_t) * d * n);
  m->bf                 = (double *)calloc((size_t)d * n, 8);
  m->R                  = (double *)calloc((size_t)d * n, 8);
  m->V                  = (double *)calloc((size_t)d * n, 8);
  m->sigma              = (double *)calloc(d, 8);
  m->hist               = (double *)calloc((size_t)d * n * n, 8);
  m->row_expo           = (int64_t *)calloc(d, 8);
  m->init_row_size      = (int *)calloc(d, sizeof(int));
  m->norm_square_b      = (double *)calloc(d, 8);
  m->expo_norm_square_b = (int64_t *)calloc(d, 8);
  for (int i = 0; i < d; i++)
  {
    int nz = n;
    while (nz > 0 && HB(i, nz - 1) == 0)
      nz--;
    m->init_row_size[i] = hmax(nz, 1);
  }
  return m;
}

void ohh_destroy(ohh_t *m)
{
  free(m->b), free(m->bf), free(m->R), free(m->V), free(m->sigma), free(m->hist);
  free(m->row_expo), free(m->init_row_size), free(m->norm_square_b), free(m->expo_norm_square_b);
  free(m);
}

static double dot_asc(const double *a, const double *b, int beg, int end)
{ /* numvect.h:385-395 */
  double r = a[beg] * b[beg];
  for (int k = beg + 1; k < end; k++)
  {
    double t = a[k] * b[k];
    r        = r + t;
  }
  return r;
}

/* refresh_R_bf(i), householder.cpp:186-245 */
void ohh_refresh_R_bf(ohh_t *m, int i)
{
  int n = m->n;
  m->n_known_cols = hmax(m->n_known_cols, m->init_row_size[i]);
  int nc          = m->n_known_cols;
  if (m->enable_row_expo)
  {
    int64_t max_expo = INT64_MIN;
    int *te          = (int *)malloc(sizeof(int) * n);
    for (int j = 0; j < nc; j++)
    {
      HBF(i, j) = frexp((double)HB(i, j), &te[j]);
      if (te[j] > max_expo)
        max_expo = te[j];
    }
    for (int j = 0; j < nc; j++)
      HBF(i, j) = ldexp(HBF(i, j), (int)(te[j] - max_expo));
    for (int j = nc; j < n; j++)
      HBF(i, j) = 0.0;
    m->row_expo[i] = max_expo;
    free(te);
  }
  else
  {
    for (int j = 0; j < nc; j++)
      HBF(i, j) = (double)HB(i, j);
    for (int j = nc; j < n; j++)
      HBF(i, j) = 0.0;
  }
  for (int j = 0; j < nc; j++)
    HR(i, j) = HBF(i, j);
  for (int j = nc; j < n; j++)
    HR(i, j) = 0.0;
  m->norm_square_b[i]      = dot_asc(&HBF(i, 0), &HBF(i, 0), 0, nc); /* norm_square_b_row, householder.h:538-551 */
  m->expo_norm_square_b[i] = m->enable_row_expo ? 2 * m->row_expo[i] : 0;
}

/* refresh_R(i), householder.cpp:247-261 */
void ohh_refresh_R(ohh_t *m, int i)
{
  for (int j = 0; j < m->n_known_cols; j++)
    HR(i, j) = HBF(i, j);
  for (int j = m->n_known_cols; j < m->n; j++)
    HR(i, j) = 0.0;
}

/* update_R_last(i), householder.cpp:27-146 */
void ohh_update_R_last(ohh_t *m, int i)
{
  int n = m->n;
  double f0, f1, f2, f3;
  m->sigma[i] = (HR(i, i) < 0) ? -1.0 : 1.0;
  if (i + 1 == n)
    f3 = 0.0;
  else
    f3 = dot_asc(&HR(i, 0), &HR(i, 0), i + 1, n);
  f1 = HR(i, i) * HR(i, i);
  f1 = f1 + f3;
  if (f1 != 0.0)
  {
    f2 = sqrt(f1);
    f0 = m->sigma[i] * f2;
    f1 = HR(i, i) + f0;
    f3 = -f3;
    f3 = f3 / f1;
    if (f3 != 0.0)
    {
      f0       = -f0;
      f0       = f0 * f3;
      f0       = sqrt(f0);
      HV(i, i) = f3 / f0;
      HR(i, i) = f2;
      for (int k = n - 1; k >= i + 1; k--) /* NumVect::div(v, b, n, c): descending, numvect.h */
        HV(i, k) = HR(i, k) / f0;
    }
    else
    {
      HV(i, i) = 0.0;
      if (HR(i, i) < 0)
        HR(i, i) = -HR(i, i);
      for (int k = i + 1; k < n; k++)
        HV(i, k) = 0.0;
    }
  }
  else
  {
    HR(i, i) = 0.0;
    HV(i, i) = 0.0;
    for (int k = i + 1; k < n; k++)
      HV(i, k) = 0.0;
  }
  m->n_known_rows++;
}

/* update_R(i, last_j), householder.cpp:151-184 */
void ohh_update_R(ohh_t *m, int i, int last_j)
{
  int n = m->n;
  if (!m->updated_R)
  {
    for (int j = 0; j < i; j++)
    {
      double f0 = dot_asc(&HV(j, 0), &HR(i, 0), j, n);
      f0        = -f0;
      for (int k = n - 1; k >= j; k--) /* NumVect::addmul(v, x, beg, n): descending, numvect.h:300-305 */
      {
        double t = HV(j, k) * f0;
        HR(i, k) = HR(i, k) + t;
      }
      HR(i, j) = m->sigma[j] * HR(i, j);
      for (int k = j; k < n; k++)
        m->hist[((size_t)i * n + j) * n + k] = HR(i, k);
    }
    if (last_j)
      ohh_update_R_last(m, i);
  }
}

static void invalidate_row(ohh_t *m, int k)
{
  if (k < m->n_known_rows)
    m->n_known_rows = k;
}

static long fexpo(double x) { return (long)ilogb(x) + 1; }

/* row_addmul_we(i, j, x, expo_add), householder.cpp:522-559 (OP_FORCE_LONG integer part + float axpy on R[i][0..i)) */
void ohh_row_addmul_we(ohh_t *m, int i, int j, double x, long expo_add)
{
  long expo;
  if (x == 0)
    expo = 0;
  else
  {
    long e = fexpo(x) + expo_add - 63;
    expo   = e > 0 ? e : 0;
  }
  long lx = (long)ldexp(x, (int)(expo_add - expo));
  int nc  = m->n_known_cols;
  for (int c = nc - 1; c >= 0; c--)
  {
    uint64_t t = (uint64_t)HB(j, c) * (uint64_t)lx;
    if (expo)
      t = expo >= 64 ? 0 : (t << expo);
    HB(i, c) = (int64_t)((uint64_t)HB(i, c) + t);
  }
  if (x == 1.0)
    for (int k = i - 1; k >= 0; k--)
      HR(i, k) = HR(i, k) + HR(j, k);
  else if (x == -1.0)
    for (int k = i - 1; k >= 0; k--)
      HR(i, k) = HR(i, k) - HR(j, k);
  else
    for (int k = i - 1; k >= 0; k--)
    {
      double t = HR(j, k) * x;
      HR(i, k) = HR(i, k) + t;
    }
}

/* size_reduce(k, size_reduction_end, size_reduction_start), householder.cpp:403-451 */
int ohh_size_reduce(ohh_t *m, int k, int sr_end, int sr_start)
{
  int reduced = 0;
  for (int i = sr_end - 1; i >= sr_start; i--)
  {
    double f  = HR(k, i) / HR(i, i);
    long de   = (long)(m->row_expo[k] - m->row_expo[i]);
    if (!(fexpo(f) + de >= 53)) /* rnd_we, nr_FP_d.inl:226-233 */
      f = ldexp(rint(ldexp(f, (int)de)), (int)-de);
    f = -f;
    if (f != 0)
    {
      ohh_row_addmul_we(m, k, i, f, de);
      reduced = 1;
    }
  }
  if (reduced)
    invalidate_row(m, k);
  return reduced;
}

static void swap_rows(void *a, void *b, size_t bytes)
{
  void *t = malloc(bytes);
  memcpy(t, a, bytes), memcpy(a, b, bytes), memcpy(b, t, bytes);
  free(t);
}

/* swap(i, j), householder.cpp:372-398 */
void ohh_swap(ohh_t *m, int i, int j)
{
  size_t n = m->n;
  invalidate_row(m, i);
  swap_rows(&HB(i, 0), &HB(j, 0), n * 8);
  swap_rows(&HBF(i, 0), &HBF(j, 0), n * 8);
  swap_rows(&m->sigma[i], &m->sigma[j], 8);
  if (m->enable_row_expo)
    swap_rows(&m->row_expo[i], &m->row_expo[j], 8);
  swap_rows(&m->init_row_size[i], &m->init_row_size[j], sizeof(int));
  swap_rows(&m->hist[i * n * n], &m->hist[j * n * n], n * n * 8);
  swap_rows(&m->norm_square_b[i], &m->norm_square_b[j], 8);
  swap_rows(&m->expo_norm_square_b[i], &m->expo_norm_square_b[j], 8);
}

/* recover_R(i), householder.h:597-608 */
void ohh_recover_R(ohh_t *m, int i)
{
  size_t n = m->n;
  for (int k = 0; k < i - 1; k++)
    HR(i, k) = m->hist[((size_t)i * n + k) * n + k];
  for (size_t k = i - 1; k < n; k++)
    HR(i, k) = m->hist[((size_t)i * n + (i - 1)) * n + k];
  m->updated_R = 1;
}

void ohh_set_updated_R_false(ohh_t *m) { m->updated_R = 0; }

/* ---- HLLLReduction<Z_NR<long>, FP_NR<double>>::hlll(), fplll/hlll.cpp:25-171 ------------------------------------
 * with size_reduction (:262-354, default branch: approx = 0.1), verify_size_reduction (:373-478, default branch:
 * the crude weak-size-reduction test) and lovasz_test (:173-236, [MSV'09] form).  compute_dR / compute_eR follow
 * hlll.h:147-159 to the letter — eR[k] is delta * R(k,k) there (not eta * R(k,k)), and so it is here.
 * Returns the RedStatus (defs.h:153-169): 0 success, 10 RED_HLLL_NORM_FAILURE, 11 RED_HLLL_SR_FAILURE. */
typedef struct
{
  double delta, eta, theta;
  double *dR, *eR;
} hlll_ctx;

static void h_compute_dR(ohh_t *m, hlll_ctx *c, int k)
{
  double f = HR(k, k);
  f        = f * f;
  c->dR[k] = c->delta * f;
}
static void h_compute_eR(ohh_t *m, hlll_ctx *c, int k) { c->eR[k] = c->delta * HR(k, k); }

static void h_size_reduction(ohh_t *m, int kappa, int sr_end, int sr_start)
{
  int not_stop = 1, prev_not_stop = 1;
  const double approx = 0.1;
  ohh_update_R(m, kappa, 0);
  m->updated_R = 0;
  for (;;)
  {
    if (!ohh_size_reduce(m, kappa, sr_end, sr_start))
      return;
    double t       = m->norm_square_b[kappa];
    int64_t expo0  = m->expo_norm_square_b[kappa];
    ohh_refresh_R_bf(m, kappa);
    double f1      = m->norm_square_b[kappa];
    int64_t expo1  = m->expo_norm_square_b[kappa];
    double f0      = approx * t;
    f0             = ldexp(f0, (int)(expo0 - expo1));
    not_stop       = (f1 <= f0);
    ohh_update_R(m, kappa, 0);
    if (prev_not_stop || not_stop)
      prev_not_stop = not_stop;
    else
      return;
  }
}

static int h_verify_size_reduction(ohh_t *m, hlll_ctx *c, int kappa)
{
  /* norm_R_row(kappa, kappa, n), householder.h:572-588 */
  double f1 = 0.0;
  if (m->n > kappa)
    f1 = sqrt(dot_asc(&HR(kappa, 0), &HR(kappa, 0), kappa, m->n));
  f1 = f1 * c->theta;
  const int64_t expo0 = m->row_expo[kappa];
  for (int i = 0; i < kappa; i++)
  {
    double f0 = fabs(HR(kappa, i));
    double f2 = ldexp(c->eR[i], (int)(m->row_expo[i] - expo0));
    f2        = f1 + f2;
    if (f0 > f2)
      return 0;
  }
  return 1;
}

static int h_lovasz_test(ohh_t *m, hlll_ctx *c, int k)
{
  double f0 = m->norm_square_b[k];
  int64_t expo1 = m->enable_row_expo ? 2 * m->row_expo[k] : 0;
  double f1 = 0.0;
  if (k - 1 > 0)
    f1 = dot_asc(&HR(k, 0), &HR(k, 0), 0, k - 1); /* norm_square_R_row(k, 0, k-1), householder.h:554-568 */
  f1 = f0 - f1;
  int64_t e0 = m->row_expo[k - 1];
  f1 = ldexp(f1, (int)(expo1 - 2 * e0));
  return c->dR[k - 1] <= f1;
}

int ohh_hlll(ohh_t *m, double delta, double eta, double theta, double cc)
{
  (void)cc; /* sr = 2^(-c d) is only read under HOUSEHOLDER_USE_SIZE_REDUCTION_TEST (hlll.cpp:296-312) */
  const int d = m->d;
  hlll_ctx c;
  c.delta = delta, c.eta = eta, c.theta = theta;
  c.dR = (double *)calloc(d, 8), c.eR = (double *)calloc(d, 8);
  double *prev_R = (double *)calloc(d, 8);
  int64_t *prev_expo = (int64_t *)calloc(d, 8);
  int status = -1;
  ohh_refresh_R_bf(m, 0);
  ohh_update_R_last(m, 0);
  h_compute_dR(m, &c, 0);
  h_compute_eR(m, &c, 0);
  int k = 1, k_max = 1, prev_k = -1;
  if (d < 2)
  {
    status = 0;
    goto done;
  }
  ohh_refresh_R_bf(m, 1);
  for (;;)
  {
    h_size_reduction(m, k, k, 0);
    if (!h_verify_size_reduction(m, &c, k))
    {
      status = 11;
      break;
    }
    if (h_lovasz_test(m, &c, k))
    {
      ohh_update_R_last(m, k);
      h_compute_dR(m, &c, k);
      h_compute_eR(m, &c, k);
      if (prev_k == k + 1)
      {
        double f0 = HR(k, k);
        double f1 = ldexp(prev_R[k], (int)(prev_expo[k] - m->row_expo[k]));
        if (f0 > f1)
        {
          status = 10;
          break;
        }
      }
      prev_k       = k;
      prev_R[k]    = HR(k, k);
      prev_expo[k] = m->row_expo[k];
      k++;
      if (k < d)
      {
        if (k > k_max)
        {
          k_max = k;
          ohh_refresh_R_bf(m, k);
        }
        else
          ohh_refresh_R(m, k);
      }
      else
      {
        status = 0;
        break;
      }
    }
    else
    {
      ohh_swap(m, k - 1, k);
      prev_k = k;
      if (k - 1 == 0)
      {
        ohh_refresh_R(m, 0);
        ohh_update_R_last(m, 0);
        h_compute_dR(m, &c, 0);
        h_compute_eR(m, &c, 0);
        ohh_refresh_R(m, 1);
        k = 1;
      }
      else
      {
        k--;
        ohh_recover_R(m, k);
      }
    }
  }
done:
  free(c.dR), free(c.eR), free(prev_R), free(prev_expo);
  return status;
}
