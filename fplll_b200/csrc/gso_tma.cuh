// gso_tma.cuh — TMA-staged variant of the batched update_gso_row sweep (gso_interface.cpp:131-164).
//
// Same algorithm, same operation order and therefore the same bits as warp_update_gso_row (gso_warp.cuh); what changes
// is how the two streamed operands reach the SM.  The register-staged version keeps ~12 x 256 B in flight per warp and
// is bound by HBM latency x bytes in flight (profiles/): registers are the limit.  Here every warp owns a ring of
// STAGES x 4 KB in shared memory that the TMA unit fills with bulk asynchronous copies (cp.async.bulk, SASS UBLKCP)
// signalled through mbarriers: 16 KB per warp in flight at no register cost.  The panel-packed layout (gso_layout.cuh)
// makes every chunk — 16 consecutive columns x 32 rows of one panel of bf or mu — one contiguous 4 KB block in HBM, so
// a chunk is exactly one bulk copy and the consumer reads it back conflict-free ([column][lane] in shared memory).
//
// Scope: full-row updates (whole Gram row invalid, gso_valid_cols[i] == 0 — what LLL's babai pays after every
// row_op_end, lll.cpp:166-224).  Panels with all 32 rows in range stream through the ring; the diagonal 32x32 tiles
// (half of their lines are never needed), a trailing partial panel and r(i,i) use the predicated-load path.
#pragma once
#include "gso_warp.cuh"

namespace b200 {

constexpr int TMA_STAGES     = 4;
constexpr int TMA_CHUNK_COLS = 16;
constexpr int TMA_CHUNK_DBL  = TMA_CHUNK_COLS * 32;  // doubles per stage (4 KB)

__device__ inline unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ inline void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ inline void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ inline void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "WAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\t"
               "bra WAIT_%=;\n\t"
               "DONE_%=:\n\t"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// global -> shared bulk copy, completion counted on the mbarrier (bytes and both addresses are multiples of 16)
__device__ inline void tma_bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// The stream of chunks of one row update: for each full panel p: the bf panel (ncols columns), then the rectangular
// part of the mu panel (32p columns).  Both producer and consumer walk it with this iterator.
struct ChunkIter
{
  int p, phase, c0;  // panel, 0 = bf / 1 = mu-rect, first column of the chunk
  int npan, ncols;
  __device__ void start(int npan_, int ncols_)
  {
    npan = npan_, ncols = ncols_, p = 0, phase = 0, c0 = 0;
  }
  __device__ bool done() const { return p >= npan; }
  __device__ int limit() const { return phase == 0 ? ncols : 32 * p; }
  __device__ int cols() const { return min(TMA_CHUNK_COLS, limit() - c0); }
  __device__ void next()
  {
    c0 += TMA_CHUNK_COLS;
    if (c0 >= limit())
    {
      c0 = 0;
      if (phase == 0 && p > 0)
        phase = 1;  // panel 0 has no rectangular part
      else
      {
        phase = 0;
        p++;
      }
    }
  }
  __device__ const double *src(const View &v) const
  {
    return phase == 0 ? v.bf + (size_t)p * 32 * v.n + (size_t)c0 * 32 : v.mu + mu_panel_base(p) + (size_t)c0 * 32;
  }
};

// ring: TMA_STAGES * TMA_CHUNK_DBL doubles (16-byte aligned), bars: TMA_STAGES mbarriers, both per warp.
// Preconditions checked by the caller: gso_valid_cols[i] == 0 and the whole Gram row i is invalid (NaN).
__device__ inline bool warp_update_gso_row_tma(const View &v, int i, int last_j, WarpSmem &s, double *ring,
                                               unsigned long long *bars, int lane)
{
  const int ncols = v.meta[M_NKC], n = v.n;
  double *gfrow = v.gf + tri_off(i), *rrow_g = v.r + tri_off(i);
  warp_stage_bf_row(v, i, ncols, s.vb, lane);
  const int jl   = min(last_j, i - 1);
  const int npan = (jl + 1) >> 5;  // panels with all 32 rows <= jl stream through the ring
  if (lane == 0)
    for (int q = 0; q < TMA_STAGES; q++)
      mbar_init(bars + q, 1);
  mbar_fence_init();
  __syncwarp();

  ChunkIter prod, cons;
  prod.start(npan, ncols);
  cons.start(npan, ncols);
  int issued = 0;
  if (lane == 0)
    for (; issued < TMA_STAGES && !prod.done(); issued++, prod.next())
    {
      const unsigned bytes = (unsigned)prod.cols() * 256u;
      mbar_expect_tx(bars + issued, bytes);
      tma_bulk_load(ring + (size_t)issued * TMA_CHUNK_DBL, prod.src(v), bytes, bars + issued);
    }
  bool ok     = true;
  int consumed = 0;
  for (int p = 0; p < npan; ++p)
  {
    const int j = 32 * p + lane;
    double acc  = 0.0;
    // ---- phase 0 (Gram, numvect.h:385-395) then phase 1 (rectangular part of the forward substitution) ----
    for (int ph = 0; ph < (p > 0 ? 2 : 1); ++ph)
    {
      const double *vec = ph == 0 ? s.vb : s.rrow;
      const int lim     = ph == 0 ? ncols : 32 * p;
      for (int c0 = 0; c0 < lim; c0 += TMA_CHUNK_COLS, consumed++, cons.next())
      {
        const int slot = consumed % TMA_STAGES;
        mbar_wait(bars + slot, (unsigned)((consumed / TMA_STAGES) & 1));
        const double *tile = ring + (size_t)slot * TMA_CHUNK_DBL + lane;
        const int nc       = min(TMA_CHUNK_COLS, lim - c0);
        if (ph == 0)
        {
#pragma unroll 4
          for (int u = 0; u < nc; u++)
          {
            const double t = __dmul_rn(tile[u * 32], vec[c0 + u]);
            acc            = (c0 + u == 0) ? t : __dadd_rn(acc, t);
          }
        }
        else
        {
#pragma unroll 4
          for (int u = 0; u < nc; u++)
            acc = __dsub_rn(acc, __dmul_rn(tile[u * 32], vec[c0 + u]));
        }
        __syncwarp();  // every lane is done with this slot before the TMA unit may overwrite it
        if (lane == 0 && !prod.done())
        {
          const unsigned bytes = (unsigned)prod.cols() * 256u;
          mbar_expect_tx(bars + slot, bytes);
          tma_bulk_load(ring + (size_t)slot * TMA_CHUNK_DBL, prod.src(v), bytes, bars + slot);
          prod.next();
        }
      }
      if (ph == 0)
        gfrow[j] = acc;  // get_gram caches the entry (gso.h:324-327)
    }
    // ---- diagonal tile: identical to warp_update_gso_row ----
    const double *mup = v.mu + mu_panel_base(p) + lane;
    double rd         = 1.0;
    {
      const double *tile = mup + (size_t)(32 * p) * 32;
      double m[8], mn[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        m[u] = (lane >= u) ? tile[(size_t)u * 32] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++)
      {
        if (q < 3)
        {
#pragma unroll
          for (int u = 0; u < 8; u++)
            mn[u] = (lane >= 8 * (q + 1) + u) ? tile[(size_t)(8 * (q + 1) + u) * 32] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
          const int t = 8 * q + u;
          if (lane == t)
            rd = m[u];
          if (t < 31)
          {
            const double rk = __shfl_sync(FULL, acc, t);
            if (lane > t)
              acc = __dsub_rn(acc, __dmul_rn(m[u], rk));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
          m[u] = mn[u];
      }
    }
    rrow_g[j]          = acc;
    s.rrow[j]          = acc;
    const double mm    = __ddiv_rn(acc, rd);
    v.mu[mu_off(i, j)] = mm;
    s.murow[j]         = mm;
    if (!isfinite(mm))
      ok = false;
    __syncwarp();
  }
  ok = __all_sync(FULL, ok);
  if (!ok)
    return false;
  // trailing partial panel (rows 32*npan .. jl) and the diagonal r(i,i): the predicated-load path, resuming at
  // column 32*npan with everything before it already in shared memory
  if (lane == 0)
    v.valid[i] = 32 * npan;
  __syncwarp();
  if (32 * npan <= last_j)
    return warp_update_gso_row(v, i, last_j, s, lane);
  return true;
}

}  // namespace b200
