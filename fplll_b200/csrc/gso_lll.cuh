// gso_lll.cuh — device-resident LLL inner loop (one warp per lattice).
//
// LLLReduction<Z_NR<long>, FP_NR<double>>::lll / babai (fplll/lll.cpp:44-224) run entirely on the device against
// the HBM-resident GSO state: no host round trip per update_gso_row / row_addmul_we (SURVEY §7 "hard parts":
// the CPU spends ~1 us per such call, less than one kernel launch).  Control flow is warp-uniform; every
// floating-point chain keeps the reference's operation order so the basis trajectory is the reference's.
#pragma once
#include "gso_warp.cuh"

namespace b200 {

enum { RED_SUCCESS = 0, RED_GSO_FAILURE = 2, RED_BABAI_FAILURE = 3, RED_LLL_FAILURE = 4 };  // defs.h:153-169
constexpr long SIZE_RED_FAILURE_THRESH = 5;                                                    // defs.h:146

struct LLLStats
{
  long n_swaps, final_kappa, zeros, babai_iters;
};

// LLLReduction::babai(kappa, size_reduction_end, size_reduction_start), lll.cpp:166-224.
// MAXQ*32 >= d.  Returns RED_SUCCESS or the failing status (warp-uniform).
template <int MAXQ>
__device__ inline int warp_babai(const View &v, WarpSmem &s, int kappa, int sr_end, int sr_start, double eta,
                                 int lane, long &iters)
{
  long max_expo = LONG_MAX;
  for (int iter = 0;; iter++)
  {
    if (!warp_update_gso_row(v, kappa, sr_end - 1, s, lane))
      return RED_GSO_FAILURE;
    // gather row kappa of mu (stride-32 in the panel layout) + exponent differences
    const int ek = v.row_expo[kappa];
    int loop_needed = 0;
    long new_max    = LONG_MIN;
    double bm[MAXQ];
#pragma unroll
    for (int q = 0; q < MAXQ; q++)
    {
      const int k = 32 * q + lane;
      bm[q]       = 0.0;
      if (k < sr_end)
      {
        bm[q]          = v.mu[mu_off(kappa, k)];
        const long de  = v.row_expo_en ? (long)(ek - v.row_expo[k]) : 0;
        s.xs[k]        = 0.0;
        if (k >= sr_start)
          loop_needed |= (fabs(ldexp(bm[q], (int)de)) > eta);  // get_mu, gso_interface.h:694-701
        new_max = max(new_max, de + fexponent(bm[q]));         // get_max_mu_exp, gso_interface.cpp:88-98
      }
    }
    if (!__any_sync(FULL, loop_needed))
      break;
    if (iter >= 2)
    {
      for (int o = 16; o; o >>= 1)
        new_max = max(new_max, __shfl_xor_sync(FULL, new_max, o));
      if (new_max > max_expo - SIZE_RED_FAILURE_THRESH)
        return RED_BABAI_FAILURE;
      max_expo = new_max;
    }
    iters++;
    __syncwarp();
    // back-substitution, j descending (lll.cpp:202-214): X_j = rnd_we(babai_mu[j]); babai_mu[k] -= X_j*mu(j,k), k<j
    for (int p = (sr_end - 1) >> 5; p >= (sr_start >> 5); --p)
    {
      // in-panel triangle: lane l owns column k = 32p+l
      double val = 0.0;
#pragma unroll
      for (int q = 0; q < MAXQ; q++)
        if (q == p)
          val = bm[q];
      const int kcol        = 32 * p + lane;
      const double *tilecol = v.mu + mu_panel_base(p) + (size_t)kcol * 32;  // mu(32p+t, kcol) at [t]
      for (int t = 31; t >= 0; --t)
      {
        const int j = 32 * p + t;
        if (j >= sr_end || j < sr_start)
          continue;
        const double bj = __shfl_sync(FULL, val, t);
        const long de   = v.row_expo_en ? (long)(ek - v.row_expo[j]) : 0;
        const double X  = rnd_we(bj, de);
        if (X == 0.0)
          continue;
        if (lane == 0)
          s.xs[j] = X;
        if (lane < t && kcol >= sr_start)
          val = __dsub_rn(val, __dmul_rn(X, tilecol[t]));
      }
      __syncwarp();
      // rectangular part: columns k < 32p (lanes over k), rows of this panel descending
#pragma unroll
      for (int q = 0; q < MAXQ; q++)
      {
        if (q < p)
        {
          const int k = 32 * q + lane;
          if (k >= sr_start)
          {
            const double *col = v.mu + mu_panel_base(p) + (size_t)k * 32;
            double a          = bm[q];
            for (int t = 31; t >= 0; --t)
            {
              const int j = 32 * p + t;
              if (j >= sr_end || j < sr_start)
                continue;
              const double X = s.xs[j];
              if (X == 0.0)
                continue;
              a = __dsub_rn(a, __dmul_rn(X, col[t]));
            }
            bm[q] = a;
          }
        }
      }
    }
    __syncwarp();
    // integer row operations b_kappa += (-X_j) * 2^expo_j * b_j, fused over j (row_addmul_we, gso.cpp:236-262).
    // Integer additions commute exactly (mod 2^64), so one pass over the columns applies all j.
    // First convert every X_j with get_si_exp_we (lanes over j): lx -> aux[j] (bit pattern), shift -> murow[j].
    for (int j = sr_start + lane; j < sr_end; j += 32)
    {
      const double X = s.xs[j];
      long expo      = 0, lx = 0;
      if (X != 0.0)
        lx = get_si_exp_we(-X, expo, v.row_expo_en ? (long)(ek - v.row_expo[j]) : 0);
      s.aux[j]   = __longlong_as_double((long long)lx);
      s.murow[j] = __longlong_as_double((long long)expo);
    }
    __syncwarp();
    {
      const int nc           = v.meta[M_NKC];
      unsigned long long *bk = (unsigned long long *)(v.b + (size_t)kappa * v.ldb);
      for (int c0 = 0; c0 < nc; c0 += 32)
      {
        const int c            = c0 + lane;
        unsigned long long acc = (c < nc) ? bk[c] : 0ull;
        for (int j = sr_end - 1; j >= sr_start; --j)
        {
          const unsigned long long lx = (unsigned long long)__double_as_longlong(s.aux[j]);
          if (lx == 0ull)
            continue;
          const long long expo = __double_as_longlong(s.murow[j]);
          if (c < nc)
          {
            unsigned long long t = ((const unsigned long long *)(v.b + (size_t)j * v.ldb))[c] * lx;
            acc += (expo >= 64 ? 0ull : (t << expo));
          }
        }
        if (c < nc)
          bk[c] = acc;
      }
    }
    __syncwarp();
    warp_row_op_end(v, kappa, kappa + 1, lane);
  }
  return RED_SUCCESS;
}

// Matrix::get_max_exp over b with Z_NR<long>::exponent (nr_Z_l.inl:40-48)
__device__ inline long warp_max_exp_of_b(const View &v, int lane)
{
  long mx = 0;
  for (int i = 0; i < v.d; i++)
    for (int c = lane; c < v.n; c += 32)
    {
      const long x = v.b[(size_t)i * v.ldb + c];
      int e;
      const double f = frexp((double)x, &e);
      long ex        = e;
      if (x > ((1L << 52) - 1) && fabs(f) == 0.5)
      {
        unsigned long long y = (unsigned long long)x;
        for (ex = 0; y; ex++, y >>= 1)
          ;
      }
      mx = max(mx, ex);
    }
  for (int o = 16; o; o >>= 1)
    mx = max(mx, __shfl_xor_sync(FULL, mx, o));
  return mx;
}

__device__ inline bool warp_b_row_is_zero(const View &v, int i, int lane)
{
  int nz = 0;
  for (int c = lane; c < v.n; c += 32)
    nz |= (v.b[(size_t)i * v.ldb + c] != 0);
  return !__any_sync(FULL, nz);
}

// get_gram(kappa,kappa) (gso.h:314-331) for the Lovasz test; computes the dot product if the entry is invalid.
__device__ inline double warp_get_gram_diag(const View &v, WarpSmem &s, int i, int lane)
{
  double *g  = v.gf + tri_off(i) + i;
  double val = *g;
  if (val != val)
  {
    const int ncols = v.meta[M_NKC];
    warp_stage_bf_row(v, i, ncols, s.vb, lane);
    __syncwarp();
    if (lane == 0)
    {
      double a = __dmul_rn(s.vb[0], s.vb[0]);
      for (int c = 1; c < ncols; c++)
        a = __dadd_rn(a, __dmul_rn(s.vb[c], s.vb[c]));
      *g = a;
    }
    __syncwarp();
    val = *g;
  }
  return val;
}

// LLLReduction::size_reduction(kappa_min, kappa_end, size_reduction_start), lll.h:106-122
template <int MAXQ>
__device__ inline int warp_size_reduction(const View &v, WarpSmem &s, int kappa_min, int kappa_end, int sr_start,
                                          double eta, int lane, long &iters)
{
  // Rows below the clean prefix are size-reduced with a valid GSO and untouched since: on them the reference's loop
  // body (babai finds nothing to reduce, update_gso_row finds the row valid) changes no state, so start after them.
  const bool eta_same = (v.meta[M_ETA_LO] == __double2loint(eta)) && (v.meta[M_ETA_HI] == __double2hiint(eta));
  const int clean     = (eta_same && sr_start == 0) ? v.meta[M_CLEAN_SR] : 0;
  const int k_first   = (kappa_min <= clean) ? max(kappa_min, min(clean, kappa_end)) : kappa_min;
  __syncwarp();
  for (int k = k_first; k < kappa_end; k++)
  {
    if (k > 0)
    {
      const int st = warp_babai<MAXQ>(v, s, k, k, sr_start, eta, lane, iters);
      if (st != RED_SUCCESS)
        return st;
    }
    if (!warp_update_gso_row(v, k, k, s, lane))
      return RED_GSO_FAILURE;  // the reference returns false here without touching status (lll.h:118-119)
  }
  if (sr_start == 0 && kappa_min <= clean && lane == 0)
  {
    // rows [0, k_first) were clean, rows [k_first, kappa_end) have just been size-reduced (row operations inside
    // this loop only lowered the marker to rows >= k_first, which were then redone)
    if (!eta_same)
    {
      v.meta[M_ETA_LO]    = __double2loint(eta);
      v.meta[M_ETA_HI]    = __double2hiint(eta);
      v.meta[M_CLEAN_LLL] = 0;
      v.meta[M_CLEAN_SR]  = kappa_end;
    }
    else
      v.meta[M_CLEAN_SR] = max(v.meta[M_CLEAN_SR], kappa_end);
  }
  __syncwarp();
  return RED_SUCCESS;
}

// LLLReduction::lll(kappa_min, kappa_start, kappa_end, size_reduction_start), lll.cpp:44-164; LLL_DEFAULT flags
// (no siegel, no early reduction, not verbose).  lov = shared array of d+1 doubles.
template <int MAXQ>
__device__ inline int warp_lll(const View &v, WarpSmem &s, double *lov, double delta, double eta, int kappa_min,
                               int kappa_start, int kappa_end, int sr_start, int lane, LLLStats &st)
{
  const int d = kappa_end - kappa_min;
  int kappa = kappa_start + 1, zeros = 0;
  st.n_swaps = st.final_kappa = st.zeros = st.babai_iters = 0;
  const double swap_threshold = delta;
  // Clean prefix: rows [0, c) are (delta, eta)-LLL-reduced with a valid GSO and untouched since the call that made
  // them so.  On such rows every iteration of the reference's loop is a no-op on the state (babai finds |mu| <= eta,
  // Lovasz holds, set_r re-writes the value the same chain produced before), so the loop may start at row c.
  const bool par_same = v.meta[M_ETA_LO] == __double2loint(eta) && v.meta[M_ETA_HI] == __double2hiint(eta) &&
                        v.meta[M_DELTA_LO] == __double2loint(delta) && v.meta[M_DELTA_HI] == __double2hiint(delta);
  const bool track    = (kappa_min == 0 && kappa_start == 0 && sr_start == 0);
  if (track && par_same)
    kappa = max(kappa, min(v.meta[M_CLEAN_LLL], kappa_end));
  __syncwarp();
  for (; zeros < d && warp_b_row_is_zero(v, 0, lane); zeros++)
    warp_move_row(v, kappa_min, kappa_end - 1 - zeros, lane);
  if (zeros < d)
  {
    if (kappa_start > 0)
    {
      const int bst = warp_babai<MAXQ>(v, s, kappa_start, kappa_start, sr_start, eta, lane, st.babai_iters);
      if (bst != RED_SUCCESS)
      {
        st.final_kappa = kappa_start, st.zeros = zeros;
        return bst;
      }
    }
    if (!warp_update_gso_row(v, kappa_start, kappa_start, s, lane))
    {
      st.final_kappa = kappa_start, st.zeros = zeros;
      return RED_GSO_FAILURE;
    }
  }
  const long maxe = warp_max_exp_of_b(v, lane);
  const long long max_iter =
      (long long)((double)d - (double)(2 * d * (d + 1)) * ((double)(maxe + 3) / log(delta)));
  long long iter;
  for (iter = 0; iter < max_iter && kappa < kappa_end - zeros; iter++)
  {
    const int bst = warp_babai<MAXQ>(v, s, kappa, kappa, sr_start, eta, lane, st.babai_iters);
    if (bst != RED_SUCCESS)
    {
      st.final_kappa = kappa, st.zeros = zeros;
      return bst;
    }
    // Lovasz test (lll.cpp:110-122): prefix chain lov[i] = lov[i-1] - mu(kappa,i-1) * r(kappa,i-1)
    const double g = warp_get_gram_diag(v, s, kappa, lane);
    for (int k = lane; k < kappa; k += 32)
      s.aux[k] = __dmul_rn(v.mu[mu_off(kappa, k)], v.r[tri_off(kappa) + k]);
    __syncwarp();
    int new_kappa = kappa, action = 0;  // 0: accept, 1: move_row(old_k,new_kappa), 2: zero vector
    if (lane == 0)
    {
      lov[0] = g;
      for (int i = 1; i <= kappa; i++)
        lov[i] = __dsub_rn(lov[i - 1], s.aux[i - 1]);
      double thr = __dmul_rn(v.r[tri_off(kappa - 1) + kappa - 1], swap_threshold);
      if (v.row_expo_en)
        thr = ldexp(thr, 2 * (v.row_expo[kappa - 1] - v.row_expo[kappa]));
      if (thr > lov[kappa - 1])
      {
        int kk = kappa;
        for (kk--; kk > kappa_min; kk--)
        {
          double t2 = __dmul_rn(v.r[tri_off(kk - 1) + kk - 1], swap_threshold);
          if (v.row_expo_en)
            t2 = ldexp(t2, 2 * (v.row_expo[kk - 1] - v.row_expo[kappa]));
          if (t2 < lov[kk - 1])
            break;
        }
        new_kappa = kk;
        action    = (lov[kk] > 0) ? 1 : 2;
      }
    }
    action    = __shfl_sync(FULL, action, 0);
    new_kappa = __shfl_sync(FULL, new_kappa, 0);
    if (action)
    {
      st.n_swaps++;
      const int old_k = kappa;
      if (action == 1)
      {
        warp_move_row(v, old_k, new_kappa, lane);
        kappa = new_kappa;
      }
      else
      {
        zeros++;
        warp_move_row(v, old_k, kappa_end - zeros, lane);
        kappa = old_k;
        continue;
      }
    }
    __syncwarp();
    warp_set_r(v, kappa, kappa, lov[kappa], lane);
    kappa++;
  }
  st.zeros = zeros;
  if (kappa < kappa_end - zeros)
    return RED_LLL_FAILURE;
  if (track && lane == 0)
  {
    const int upto = kappa_end - zeros;
    if (!par_same)
    {
      v.meta[M_ETA_LO] = __double2loint(eta), v.meta[M_ETA_HI] = __double2hiint(eta);
      v.meta[M_DELTA_LO] = __double2loint(delta), v.meta[M_DELTA_HI] = __double2hiint(delta);
      v.meta[M_CLEAN_LLL] = upto;
      v.meta[M_CLEAN_SR]  = upto;
    }
    else
    {
      v.meta[M_CLEAN_LLL] = max(v.meta[M_CLEAN_LLL], upto);
      v.meta[M_CLEAN_SR]  = max(v.meta[M_CLEAN_SR], upto);
    }
  }
  __syncwarp();
  return RED_SUCCESS;
}

}  // namespace b200
