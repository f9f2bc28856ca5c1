// enum_api.cu — B200-native Schnorr-Euchner enumeration (BKZ's SVP subtree search) behind include/b200enum.h.
//
// Reference algorithm: EnumerationBase::enumerate_loop (fplll/enum/enumerate_base.cpp:152-254) + next_pos_up
// (enumerate_base.h:145-171); the only parallel strategy in the reference is enumlib's subtree fan-out over
// std::threads (fplll/enum-parallel/enumeration.h:382-510).  B200 design:
//   1. HOST BREADTH PHASE  — the top T levels (d-1 .. d-T) are expanded in exact Schnorr-Euchner order into subtree
//      roots (x[d-T..d-1], partial length).  T grows until there are enough roots to occupy the machine.  Roots are
//      sorted by partial length so the most promising subtrees run first (enumlib does the same, enumeration.h:417-422).
//   2. DEVICE DEPTH PHASE  — one THREAD per subtree root, roots handed out by an atomic ticket; each thread runs the
//      same iterative walk over levels d-T-1 .. 0 with its coefficient/centre/partial-length stacks in local memory
//      and mu^T, r_ii, pruning in shared memory.  148 SMs x 512 resident walkers; fp64 throughout (B200's 37 TFLOP/s
//      fp64 is what makes "recompute the centre chain on every descent" cheaper than enumlib's d x d partial-sum cache,
//      which would not fit per thread).
//   3. RADIUS — one 8-byte word in global memory, lowered with atomicMin on the bit pattern of the (positive) squared
//      length; walkers re-read it between subtrees and every 64 steps.  This is FastEvaluator "best 1" semantics
//      (enum/evaluator.h:122-156), which is what BKZ uses (bkz.h:324).
// Arithmetic: every centre is the chain ((0 - x[d-1] mu) - x[d-2] mu) - ... in descending j with separately rounded
// multiply and subtract (--fmad=false), the order of the reference's center_partsums, so with a fixed radius the set
// of visited nodes — and therefore the node count — is identical to the reference's own enumerator.
#include "../../include/b200enum.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <mutex>
#include <string>
#include <vector>
#include <deque>

namespace {

thread_local std::string g_err;
#define CKE(call)                                                                                  \
  do                                                                                               \
  {                                                                                                \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
    {                                                                                              \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
      return B200ENUM_ECUDA;                                                                       \
    }                                                                                              \
  } while (0)

constexpr int SOL_CAP    = 4096;
constexpr int THREADS    = 128;
constexpr int THREADS_XS = 512;      // upper bound of the CTA size of the x-in-shared-memory variant (one CTA per SM)
constexpr int MIN_ROOTS  = 256;      // host breadth phase: grow T until at least this many roots (the device multiplies
                                    // them by work splitting, a round costs one grid barrier) ...
constexpr int MAX_ROOTS  = 1 << 18;  // ... but never beyond this
constexpr int SMEM_XS_MAX = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100
constexpr unsigned TASK_CAP = 1u << 21;  // device task queue capacity (tasks of all rounds)

struct SolRec
{
  double dist;
  int x[B200ENUM_MAX_DIM];
};

// A unit of work: "at level lvl, with x[lvl+1..d-1] fixed, walk the remaining siblings starting at x[lvl] = xs (in
// Schnorr-Euchner order around cen) and everything below them".  The host's subtree roots are tasks; a walker that
// has used up its node budget turns the unvisited siblings of every ancestor on its path into new tasks (the classic
// depth-first work split), so the heavy-tailed subtree sizes of pruned enumeration get re-balanced between rounds.
struct TaskHdr
{
  double pd, cen, xs;
  int lvl, pad;
};

struct EnumArgs
{
  int d, dstride;              // dim; task prefix row stride (d rounded up to a multiple of 4 ints)
  const double *mut, *rdiag, *prun;
  TaskHdr *hdrq[2];            // double-buffered task queue: round r reads half r&1, appends to the other half
  int *txq[2];                 // [TASK_CAP * dstride] coefficients by absolute level (entries > lvl are meaningful)
  unsigned n_first;            // number of tasks of round 0 (the host's subtree roots, in half 0)
  unsigned *ctr;               // device counters: [0] ticket, [1] append position, [2] tasks of the current round
  unsigned yield_nodes;          // a walker re-checks the split / yield conditions every this many nodes
  unsigned budget0, budget_mul;  // nodes a walker may visit before it must split: budget0 * mul^round (capped)
  unsigned long long *A_bits;  // [0] radius (bit pattern of a positive double), [1] best-so-far in fixed-radius mode
  unsigned long long *nodes;   // [d]
  unsigned long long *leaves;
  unsigned *sol_count;
  SolRec *sols;
  int fixed_radius;
};

__device__ inline double next_sibling(double x, double c, double pdk)
{
  // next_pos_up's coefficient update (enumerate_base.h:145-171) with the zig-zag state (dx, ddx) re-derived from
  // (x, centre): x0 = round(c), s = +1 if c >= x0 else -1, sequence x0, x0+s, x0-s, x0+2s, ...
  if (pdk == 0.0)
    return x + 1.0;  // SVP: only the positive half at the top non-zero coefficient
  const double x0 = round(c), s = (c >= x0) ? 1.0 : -1.0, t = x - x0;
  return (t == 0.0) ? x0 + s : ((t * s > 0.0) ? x0 - t : x0 - t + s);
}

// ------------------------------------------------------------------------------------------------------------
// device depth phase
// Persistent cooperative kernel: ALL rounds of one enumeration run inside one launch, separated by grid-wide barriers
// (a round = every walker works off the current half of the task queue, walkers that split or yield append to the other
// half).  One launch per Enumeration::enumerate call instead of one launch + host synchronisation per round.
//
// XS = true (dim <= 64, every BKZ block size in use): the coefficient vectors x[] of all walkers of the CTA live in
// SHARED memory ([level][thread], conflict-free) — the centre chain reads x[j] d-k times per node, and with x[] in
// thread-local memory that traffic (L1 misses for 3/4 of it, ~1.6 KB of L2 reads per node) was what bounded the first
// version (profiles/r1_enum_ncu.txt).  One CTA per SM then shares a single copy of mu^T.
// Both variants also keep pre[k] = the part of level k's centre chain that only involves the task's FIXED coefficients
// (j > top0): it is computed the first time the walker reaches level k inside a task and is the exact prefix of the
// reference's descending chain, so the remaining chain is top0-k long instead of d-1-k.
template <int ML, bool XS>
__global__ void __launch_bounds__(XS ? THREADS_XS : THREADS) k_enum(EnumArgs a)
{
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double sm[];
  const int d  = a.d;
  const int ds = XS ? (d | 1) : d;  // odd row stride: walkers on different levels read different rows of mu^T
  double *s_mut = sm, *s_r = sm + (size_t)d * ds, *s_p = s_r + d;
  double *s_x = s_p + d + threadIdx.x;  // XS: x[j] of this thread at s_x[j * blockDim.x]
  const int xstr = blockDim.x;
  for (int t = threadIdx.x; t < d * d; t += blockDim.x)
    s_mut[(t / d) * ds + (t % d)] = a.mut[t];
  for (int t = threadIdx.x; t < d; t += blockDim.x)
  {
    s_r[t] = a.rdiag[t];
    s_p[t] = a.prun[t];
  }
  __syncthreads();

  double xl[XS ? 1 : ML], cen[ML], pd[ML], pre[ML];
  unsigned cnt[ML];
  auto getx = [&](int j) -> double { return XS ? s_x[(size_t)j * xstr] : xl[XS ? 0 : j]; };
  auto setx = [&](int j, double val) {
    if (XS)
      s_x[(size_t)j * xstr] = val;
    else
      xl[XS ? 0 : j] = val;
  };
  int top0 = 0, pvalid = 0;
#pragma unroll 1
  for (int k = 0; k < d; k++)
    cnt[k] = 0;
  unsigned long long my_leaves = 0;
  double A     = __longlong_as_double(*(volatile unsigned long long *)a.A_bits);
  int steps    = 0;
  unsigned end = a.n_first, budget = a.budget0;

  for (unsigned rno = 0;; ++rno)
  {
  const TaskHdr *hdr_in = a.hdrq[rno & 1];
  const int *tx_in      = a.txq[rno & 1];
  TaskHdr *hdr_out      = a.hdrq[(rno + 1) & 1];
  int *tx_out           = a.txq[(rno + 1) & 1];
  unsigned *ticket = a.ctr, *tail = a.ctr + 1;
  const unsigned total_warps  = gridDim.x * (blockDim.x >> 5);
  const unsigned lanes_allowed = min(32u, max(1u, (end + total_warps - 1) / total_warps));
  int k        = -2;  // -2: idle (needs a task)
  int top      = 0;   // highest level this walker still owns
  unsigned n   = 0;   // nodes since the last split
  unsigned next_check = a.yield_nodes;
  for (;;)
  {
    if (k == -2)
    {
      // a round with fewer tasks than lanes is spread over WARPS first: 32 walkers of one warp sit on different
      // levels and serialise each other (SIMT divergence), a lone walker in a warp runs at full single-thread speed
      if ((threadIdx.x & 31u) >= lanes_allowed)
        break;
      const unsigned t = atomicAdd(ticket, 1u);
      if (t >= end)
        break;
      const TaskHdr h = hdr_in[t];
      const int4 *tx4 = (const int4 *)(tx_in + (size_t)t * a.dstride);
      A               = __longlong_as_double(*(volatile unsigned long long *)a.A_bits);
      top = k = h.lvl;
      top0 = pvalid = k;
      // prefix x[lvl+1 .. d-1]: 128-bit loads, 4 in flight (a walker that starts a task stalls its whole warp, so
      // this must cost one memory latency, not d of them)
#pragma unroll 4
      for (int j4 = (k + 1) >> 2; j4 < (a.dstride >> 2); j4++)
      {
        const int4 q = tx4[j4];
        const int j  = 4 * j4;
        if (j > k && j < d)
          setx(j, (double)q.x);
        if (j + 1 > k && j + 1 < d)
          setx(j + 1, (double)q.y);
        if (j + 2 > k && j + 2 < d)
          setx(j + 2, (double)q.z);
        if (j + 3 > k && j + 3 < d)
          setx(j + 3, (double)q.w);
      }
      pd[k]  = h.pd;
      cen[k] = h.cen;
      setx(k, h.xs);
      n      = 0;
      next_check = a.yield_nodes;
    }

    // ---- one step of enumerate_loop (enumerate_base.cpp:193-254) ----
    const double xk      = getx(k);
    const double alphak  = __dsub_rn(xk, cen[k]);
    const double newdist = __dadd_rn(pd[k], __dmul_rn(__dmul_rn(alphak, alphak), s_r[k]));
    bool up              = true;
    if (newdist <= __dmul_rn(s_p[k], A))
    {
      cnt[k]++;
      n++;
      if (k == 0)
      {
        if (newdist > 0.0)
        {
          my_leaves++;
          const unsigned long long nb  = (unsigned long long)__double_as_longlong(newdist);
          const unsigned long long old = atomicMin(a.A_bits + (a.fixed_radius ? 1 : 0), nb);
          if (nb < old)
          {
            const unsigned slot = atomicAdd(a.sol_count, 1u);
            if (slot < SOL_CAP)
            {
              SolRec *s = a.sols + slot;
              s->dist   = newdist;
#pragma unroll 1
              for (int j = 0; j < d; j++)
                s->x[j] = (int)getx(j);
            }
          }
          if (!a.fixed_radius)
            A = fmin(A, newdist);
        }
        k = -1;  // the reference decrements first and lets next_pos_up come back to level 0
      }
      else
      {
        --k;
        // centre: the reference's chain ((0 - x[d-1] mu) - x[d-2] mu) - ... - x[k+1] mu, descending j
        double nc;
        const double *mrow = s_mut + (size_t)k * ds;
        if (k < pvalid)
        {
          // first visit of level k inside this task (k == pvalid - 1): the fixed part of the chain, j = d-1 .. top0+1
          nc = 0.0;
#pragma unroll 4
          for (int j = d - 1; j > top0; --j)
            nc = __dsub_rn(nc, __dmul_rn(getx(j), mrow[j]));
          pre[k] = nc;
          pvalid = k;
        }
        else
          nc = pre[k];
#pragma unroll 4
        for (int j = top0; j > k; --j)
          nc = __dsub_rn(nc, __dmul_rn(getx(j), mrow[j]));
        cen[k] = nc;
        pd[k]  = newdist;
        setx(k, round(nc));
        up     = false;
      }
    }
    if (up)
    {
      ++k;
      if (k > top)
        k = -2;  // everything this walker owned is done
      else
        setx(k, next_sibling(getx(k), cen[k], pd[k]));
    }
    else if (n >= next_check)
    {
      // Work split (we are at a just-created node: level k holds its first, not yet tested, candidate x[k]).
      //  * the round still has unclaimed tasks: a walker that used up its budget donates the unvisited siblings of
      //    its two top-most ancestor levels (the biggest chunks) and goes on;
      //  * the round has run dry (idle lanes are waiting): after only 64 more nodes the walker YIELDS — every
      //    ancestor's remaining siblings plus its current position become tasks of the next round — so the tail of
      //    a round is bounded by ~64 nodes instead of by the largest subtree.
      const bool dry = (*(volatile unsigned *)ticket >= end);
      next_check     = n + a.yield_nodes;
      if (dry || n >= budget)
      {
        int jj = top, given = 0;
        bool full = false;
        for (; jj > k && (dry || given < 2); --jj)
        {
          // siblings come in order of increasing distance from the centre: if the next one is already outside the
          // bound there is nothing left at this level
          const double nx = next_sibling(getx(jj), cen[jj], pd[jj]);
          const double al = __dsub_rn(nx, cen[jj]);
          const double nd = __dadd_rn(pd[jj], __dmul_rn(__dmul_rn(al, al), s_r[jj]));
          if (!(nd <= __dmul_rn(s_p[jj], A)))
            continue;
          const unsigned slot = atomicAdd(tail, 1u);
          if (slot >= TASK_CAP)
          {
            full = true;
            break;  // queue full: keep levels <= jj ourselves
          }
          TaskHdr h;
          h.pd = pd[jj], h.cen = cen[jj], h.xs = nx, h.lvl = jj, h.pad = 0;
          hdr_out[slot] = h;
          int *tx       = tx_out + (size_t)slot * a.dstride;
#pragma unroll 4
          for (int j = jj + 1; j < d; j++)
            tx[j] = (int)getx(j);
          given++;
        }
        top = jj;  // levels above jj are donated or exhausted; the walker keeps the siblings of levels <= jj
        if (dry && !full && top == k)
        {
          const unsigned slot = atomicAdd(tail, 1u);
          if (slot < TASK_CAP)
          {
            TaskHdr h;
            h.pd = pd[k], h.cen = cen[k], h.xs = getx(k), h.lvl = k, h.pad = 0;
            hdr_out[slot] = h;
            int *tx       = tx_out + (size_t)slot * a.dstride;
#pragma unroll 4
            for (int j = k + 1; j < d; j++)
              tx[j] = (int)getx(j);
            k = -2;  // yielded: this walker is idle (and the round is dry, so it will leave the loop)
          }
        }
        n          = 0;
        next_check = a.yield_nodes;
      }
    }
    if (((++steps) & 63) == 0 && !a.fixed_radius)
      A = __longlong_as_double(*(volatile unsigned long long *)a.A_bits);
  }
  // ---- end of round: everybody has left the walker loop; publish the next round's task count ----
  __threadfence();
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    const unsigned produced = *(volatile unsigned *)tail;
    a.ctr[2] = produced < TASK_CAP ? produced : TASK_CAP;
    a.ctr[0] = 0;
    a.ctr[1] = 0;
    a.ctr[3] = rno + 1;
    __threadfence();
  }
  grid.sync();
  end = *(volatile unsigned *)(a.ctr + 2);
  if (end == 0)
    break;
  const unsigned long long nb = (unsigned long long)budget * a.budget_mul;
  budget = nb > 16384ull ? 16384u : (unsigned)nb;
  }
#pragma unroll 1
  for (int kk = 0; kk < d; kk++)
    if (cnt[kk])
      atomicAdd(a.nodes + kk, (unsigned long long)cnt[kk]);
  if (my_leaves)
    atomicAdd(a.leaves, my_leaves);
}

// ------------------------------------------------------------------------------------------------------------
// host breadth phase: exact Schnorr-Euchner order over levels [L, d)
struct Breadth
{
  std::vector<int> rootx;        // nroots * T
  std::vector<double> rootdist;  // nroots
  std::vector<uint64_t> nodes;   // d
};

bool expand_top(int d, int L, const double *mut, const double *rdiag, const double *prun, double A, size_t cap,
                Breadth &out)
{
  const int T = d - L;
  std::vector<double> x(d + 1, 0.0), cen(d + 1, 0.0), pd(d + 2, 0.0);
  std::vector<int> dx(d + 1, 1), ddx(d + 1, 1);
  out.rootx.clear(), out.rootdist.clear();
  out.nodes.assign(d, 0);
  auto emit = [&](double dist) {
    for (int t = 0; t < T; t++)
      out.rootx.push_back((int)x[L + t]);
    out.rootdist.push_back(dist);
  };
  emit(0.0);  // the all-zero prefix (never counted, see k_enum)
  int k = L - 1;
  for (;;)
  {
    // next_pos_up
    ++k;
    if (pd[k] != 0.0)
    {
      x[k] += dx[k];
      ddx[k] = -ddx[k];
      dx[k]  = ddx[k] - dx[k];
    }
    else
    {
      if (k >= d)
        break;
      x[k] += 1.0;
    }
    // descend while the bound holds
    for (;;)
    {
      const double alphak  = x[k] - cen[k];
      const double newdist = pd[k] + alphak * alphak * rdiag[k];
      if (!(newdist <= prun[k] * A))
        break;
      out.nodes[k]++;
      if (k == L)
      {
        emit(newdist);
        if (out.rootdist.size() > cap)
          return false;
        k = L - 1;  // the subtree below is the device's; resume as if it had been exhausted
        break;
      }
      --k;
      double nc = 0.0;
      for (int j = d - 1; j > k; --j)
        nc = nc - x[j] * mut[(size_t)k * d + j];
      cen[k] = nc;
      pd[k]  = newdist;
      x[k]   = std::round(nc);
      dx[k] = ddx[k] = (nc >= x[k]) ? 1 : -1;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------------------
// per-device persistent buffers
struct DevCtx
{
  int device = -1, sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double *d_cfg = nullptr;  // mut | rdiag | prun
  size_t cfg_cap = 0;
  TaskHdr *d_hdr = nullptr;
  int *d_tx = nullptr;
  size_t tx_cap = 0;                      // ints
  unsigned long long *d_words = nullptr;  // [0]=A_bits [1]=best bits (fixed mode) [2]=leaves [3]=ticket|sol_count [4]=tail
  unsigned long long *d_nodes = nullptr;
  SolRec *d_sols = nullptr;
  SolRec *h_sols = nullptr;  // pinned
  unsigned long long *h_words = nullptr, *h_nodes = nullptr;
};
std::mutex g_mu;
std::deque<DevCtx> g_ctx;  // deque: get_ctx hands out pointers that must survive later push_backs

int get_ctx(int device, DevCtx **out)
{
  for (auto &c : g_ctx)
    if (c.device == device)
    {
      *out = &c;
      return 0;
    }
  DevCtx c;
  c.device = device;
  CKE(cudaSetDevice(device));
  CKE(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, device));
  CKE(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CKE(cudaEventCreate(&c.e0));
  CKE(cudaEventCreate(&c.e1));
  CKE(cudaMalloc(&c.d_words, 8 * sizeof(unsigned long long)));
  CKE(cudaMalloc(&c.d_nodes, B200ENUM_MAX_DIM * sizeof(unsigned long long)));
  CKE(cudaMalloc(&c.d_sols, SOL_CAP * sizeof(SolRec)));
  CKE(cudaMalloc(&c.d_hdr, (size_t)2 * TASK_CAP * sizeof(TaskHdr)));
  CKE(cudaMallocHost(&c.h_sols, SOL_CAP * sizeof(SolRec)));
  CKE(cudaMallocHost(&c.h_words, 8 * sizeof(unsigned long long)));
  CKE(cudaMallocHost(&c.h_nodes, B200ENUM_MAX_DIM * sizeof(unsigned long long)));
  CKE(cudaFuncSetAttribute((const void *)k_enum<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CKE(cudaFuncSetAttribute((const void *)k_enum<160, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CKE(cudaFuncSetAttribute((const void *)k_enum<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_XS_MAX));
  g_ctx.push_back(c);
  *out = &g_ctx.back();
  return 0;
}

template <class T> int ensure(T **p, size_t *cap, size_t need)
{
  if (*cap >= need)
    return 0;
  if (*p)
    cudaFree(*p);
  CKE(cudaMalloc(p, need * sizeof(T)));
  *cap = need;
  return 0;
}

}  // namespace

extern "C" {

const char *b200enum_last_error(void) { return g_err.c_str(); }

int b200enum_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b200enum_run(int dim, double maxdist, const double *mut, const double *rdiag, const double *pruning, int flags,
                 const int *devices, int ndev, int shard_rank, int shard_world, b200enum_sol_cb cb, void *ctx,
                 uint64_t *nodes, b200enum_stats *stats)
{
  if (flags & (B200ENUM_DUAL | B200ENUM_FINDSUBSOLS))
    return B200ENUM_UNSUPPORTED;
  if (dim < 2 || dim > B200ENUM_MAX_DIM || !mut || !rdiag || !(maxdist > 0) || shard_world < 1 || shard_rank < 0 ||
      shard_rank >= shard_world)
  {
    g_err = "b200enum_run: bad arguments";
    return B200ENUM_EINVAL;
  }
  const int navail = b200enum_device_count();
  if (navail == 0)
  {
    g_err = "b200enum_run: no CUDA device (this library has no CPU fallback)";
    return B200ENUM_ENODEV;
  }
  int dev0 = 0;
  if (!devices || ndev <= 0)
  {
    devices = &dev0;
    ndev    = 1;
  }
  for (int i = 0; i < ndev; i++)
    if (devices[i] < 0 || devices[i] >= navail)
    {
      g_err = "b200enum_run: bad device ordinal";
      return B200ENUM_EINVAL;
    }
  std::lock_guard<std::mutex> lock(g_mu);
  const auto t_begin = std::chrono::steady_clock::now();
  const int d = dim;
  std::vector<double> prun(d, 1.0);
  if (pruning)
    std::copy(pruning, pruning + d, prun.begin());
  const bool fixed = (flags & B200ENUM_FIXED_RADIUS) != 0;

  // ---- host breadth phase: grow T until there are enough roots ----
  Breadth br;
  bool have = false;
  for (int T = 1; T <= d - 1; ++T)
  {
    Breadth cur;
    if (!expand_top(d, d - T, mut, rdiag, prun.data(), maxdist, MAX_ROOTS, cur))
      break;  // too many: keep the previous depth
    br   = std::move(cur);
    have = true;
    if ((int)br.rootdist.size() >= MIN_ROOTS * ndev * shard_world)
      break;
  }
  if (!have)
  {
    g_err = "b200enum_run: breadth phase overflow at the first level";
    return B200ENUM_EINVAL;
  }
  const int T = (int)(br.rootx.size() / br.rootdist.size());  // depth of the expansion that was kept
  const int L = d - T;
  const size_t nroots = br.rootdist.size();

  // roots sorted by partial length (ascending; the all-zero prefix first): most promising subtrees first
  std::vector<unsigned> order(nroots);
  for (size_t i = 0; i < nroots; i++)
    order[i] = (unsigned)i;
  std::stable_sort(order.begin(), order.end(),
                   [&](unsigned a, unsigned b) { return br.rootdist[a] < br.rootdist[b]; });

  const auto t_host = std::chrono::steady_clock::now();
  // ---- device depth phase ----
  const int dstride  = (d + 3) & ~3;
  const size_t cfg_n = (size_t)d * d + 2 * d;
  std::vector<double> cfg(cfg_n);
  std::copy(mut, mut + (size_t)d * d, cfg.begin());
  std::copy(rdiag, rdiag + d, cfg.begin() + (size_t)d * d);
  std::copy(prun.begin(), prun.end(), cfg.begin() + (size_t)d * d + d);
  // this shard: sorted position g with g % shard_world == shard_rank; among those, device q takes every ndev-th
  std::vector<DevCtx *> ctxs(ndev);
  std::vector<unsigned> tail(ndev, 0);
  for (int q = 0; q < ndev; q++)
  {
    DevCtx *c;
    int rc = get_ctx(devices[q], &c);
    if (rc)
      return rc;
    ctxs[q] = c;
    CKE(cudaSetDevice(c->device));
    rc = ensure(&c->d_cfg, &c->cfg_cap, cfg_n);
    // sized for dim <= 64 up front (BKZ calls with every block size from 2 to beta: growing would re-allocate ~1 GB
    // a dozen times), re-allocated once if a larger dimension ever shows up
    rc |= ensure(&c->d_tx, &c->tx_cap, (size_t)2 * TASK_CAP * (dstride <= 64 ? 64 : B200ENUM_MAX_DIM));
    if (rc)
      return rc;
    std::vector<TaskHdr> hdr;
    std::vector<int> tx;
    for (size_t g = (size_t)shard_rank + (size_t)q * shard_world; g < nroots; g += (size_t)shard_world * ndev)
    {
      const unsigned r = order[g];
      const int *rx    = &br.rootx[(size_t)r * T];
      TaskHdr h;
      h.lvl = L - 1, h.pad = 0, h.pd = br.rootdist[r];
      double nc = 0.0;  // centre of level L-1 under this prefix (same chain as the device)
      for (int j = d - 1; j >= L; --j)
        nc = nc - (double)rx[j - L] * mut[(size_t)(L - 1) * d + j];
      h.cen = nc, h.xs = std::round(nc);
      hdr.push_back(h);
      const size_t o = tx.size();
      tx.resize(o + dstride, 0);
      for (int t = 0; t < T; t++)
        tx[o + L + t] = rx[t];
    }
    tail[q] = (unsigned)hdr.size();
    CKE(cudaMemcpyAsync(c->d_cfg, cfg.data(), cfg_n * 8, cudaMemcpyHostToDevice, c->stream));
    if (!hdr.empty())
    {
      CKE(cudaMemcpyAsync(c->d_hdr, hdr.data(), hdr.size() * sizeof(TaskHdr), cudaMemcpyHostToDevice, c->stream));
      CKE(cudaMemcpyAsync(c->d_tx, tx.data(), tx.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    unsigned long long w[8] = {0};
    memcpy(&w[0], &maxdist, 8);
    w[1] = ~0ull;  // best-so-far bits in fixed-radius mode
    CKE(cudaMemcpyAsync(c->d_words, w, sizeof(w), cudaMemcpyHostToDevice, c->stream));
    CKE(cudaMemsetAsync(c->d_nodes, 0, B200ENUM_MAX_DIM * sizeof(unsigned long long), c->stream));
    CKE(cudaStreamSynchronize(c->stream));  // hdr/tx are stack vectors
    CKE(cudaEventRecord(c->e0, c->stream));
  }
  // One cooperative launch per device runs every round of the enumeration (k_enum): walkers that exhaust their node
  // budget, or notice that the round has run dry, append the unvisited parts of their subtree as tasks of the next
  // round.  The budget grows geometrically: short first rounds multiply the parallelism, later rounds amortise the
  // grid barrier.
  static const unsigned budget0 = getenv("B200_ENUM_BUDGET0") ? atoi(getenv("B200_ENUM_BUDGET0")) : 64;
  static const unsigned budget_mul = getenv("B200_ENUM_BUDGET_MUL") ? atoi(getenv("B200_ENUM_BUDGET_MUL")) : 4;
  static const unsigned yield_nodes = getenv("B200_ENUM_YIELD") ? atoi(getenv("B200_ENUM_YIELD")) : 64;
  static const int bpsm = getenv("B200_ENUM_BLOCKS_PER_SM") ? atoi(getenv("B200_ENUM_BLOCKS_PER_SM")) : 4;
  static const int use_xs = getenv("B200_ENUM_XS") ? atoi(getenv("B200_ENUM_XS")) : 1;
  static const int xs_threads_cap = getenv("B200_ENUM_XS_THREADS") ? atoi(getenv("B200_ENUM_XS_THREADS")) : 320;
  const bool xs = use_xs && d <= 64;
  // x-in-shared variant: one CTA per SM, as many walkers as fit next to mu^T (odd row stride) + rdiag + pruning
  const size_t cfg_smem = xs ? ((size_t)d * (d | 1) + 2 * d) * sizeof(double) : cfg_n * sizeof(double);
  int threads = THREADS;
  if (xs)
  {
    threads = (int)((SMEM_XS_MAX - cfg_smem) / ((size_t)d * sizeof(double)));
    threads = std::min(std::min(threads, xs_threads_cap), THREADS_XS) & ~31;
    if (threads < 32)
      threads = 32;
  }
  const size_t smem = cfg_smem + (xs ? (size_t)threads * d * sizeof(double) : 0);
  for (int q = 0; q < ndev; q++)
  {
    if (tail[q] == 0)
      continue;
    DevCtx *c = ctxs[q];
    CKE(cudaSetDevice(c->device));
    EnumArgs a;
    a.d = d, a.dstride = dstride, a.mut = c->d_cfg, a.rdiag = c->d_cfg + (size_t)d * d, a.prun = a.rdiag + d;
    a.hdrq[0] = c->d_hdr, a.hdrq[1] = c->d_hdr + TASK_CAP;
    a.txq[0] = c->d_tx, a.txq[1] = c->d_tx + (size_t)TASK_CAP * dstride;
    a.n_first = tail[q];
    a.ctr = (unsigned *)(c->d_words + 4), a.sol_count = (unsigned *)(c->d_words + 3) + 1;
    a.budget0 = budget0, a.budget_mul = budget_mul, a.yield_nodes = yield_nodes;
    a.A_bits = c->d_words, a.leaves = c->d_words + 2, a.nodes = c->d_nodes, a.sols = c->d_sols;
    a.fixed_radius = fixed ? 1 : 0;
    int occ = 0;
    const void *fn = xs ? (const void *)k_enum<64, true>
                        : (d <= 64 ? (const void *)k_enum<64, false> : (const void *)k_enum<160, false>);
    CKE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
    if (occ < 1)
    {
      g_err = "b200enum_run: kernel does not fit on an SM";
      return B200ENUM_ECUDA;
    }
    const int blocks = c->sms * (xs ? 1 : std::min(occ, bpsm));  // every CTA must be resident (grid-wide barrier)
    void *params[]   = {(void *)&a};
    CKE(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), params, smem, c->stream));
  }
  int rounds = 0;
  std::vector<uint64_t> tot(d, 0);
  uint64_t host_nodes = 0, dev_nodes = 0, leaves = 0;
  if (shard_rank == 0)
    for (int k = 0; k < d; k++)
    {
      tot[k] += br.nodes[k];
      host_nodes += br.nodes[k];
    }
  std::vector<SolRec> found;
  float ms_max  = 0;
  bool overflow = false;
  for (int q = 0; q < ndev; q++)
  {
    DevCtx *c = ctxs[q];
    CKE(cudaSetDevice(c->device));
    CKE(cudaEventRecord(c->e1, c->stream));
    CKE(cudaMemcpyAsync(c->h_words, c->d_words, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaMemcpyAsync(c->h_nodes, c->d_nodes, d * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaStreamSynchronize(c->stream));
    const unsigned nsol = ((unsigned *)(c->h_words + 3))[1];
    if (nsol)
    {
      CKE(cudaMemcpyAsync(c->h_sols, c->d_sols, std::min<unsigned>(nsol, SOL_CAP) * sizeof(SolRec),
                          cudaMemcpyDeviceToHost, c->stream));
      CKE(cudaStreamSynchronize(c->stream));
    }
    CKE(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, c->e0, c->e1);
    ms_max = std::max(ms_max, ms);
    for (int k = 0; k < d; k++)
    {
      tot[k] += c->h_nodes[k];
      dev_nodes += c->h_nodes[k];
    }
    leaves += c->h_words[2];
    rounds = std::max(rounds, (int)((unsigned *)(c->h_words + 4))[3]);
    if (nsol > SOL_CAP)
      overflow = true;
    for (unsigned s = 0; s < std::min<unsigned>(nsol, SOL_CAP); s++)
      found.push_back(c->h_sols[s]);
  }
  // the all-zero prefix is walked by the device from level L-1 down; the reference does not count it on levels >= 1
  // (initial-descent compensation, enumerate_base.cpp:165-183) — only its level-0 node
  if (shard_rank == 0)
    for (int k = 1; k < L; k++)
    {
      tot[k]--;
      dev_nodes--;
    }
  // replay the improving solutions in order of improvement through the evaluator callback
  std::sort(found.begin(), found.end(), [](const SolRec &a, const SolRec &b) { return a.dist > b.dist; });
  double cur = maxdist;
  int nrep   = 0;
  std::vector<double> sol(d);
  if (fixed)
  {
    if (!found.empty() && cb)
    {
      const SolRec &s = found.back();
      for (int j = 0; j < d; j++)
        sol[j] = s.x[j];
      cb(ctx, s.dist, sol.data());
      nrep = 1;
    }
  }
  else
    for (const SolRec &s : found)
    {
      if (!(s.dist < cur))
        continue;
      for (int j = 0; j < d; j++)
        sol[j] = s.x[j];
      cur = cb ? cb(ctx, s.dist, sol.data()) : s.dist;
      nrep++;
    }
  if (nodes)
    for (int k = 0; k < d; k++)
      nodes[k] = tot[k];
  if (stats)
  {
    stats->host_nodes = host_nodes, stats->device_nodes = dev_nodes, stats->leaves = leaves;
    stats->top_levels = T, stats->n_roots = (int)nroots, stats->n_solutions = nrep, stats->n_devices = ndev;
    stats->n_rounds = rounds;
    stats->final_maxdist = cur, stats->device_ms = ms_max;
    stats->host_breadth_us = (float)std::chrono::duration<double, std::micro>(t_host - t_begin).count();
    stats->total_us = (float)std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count();
  }
  if (overflow)
  {
    g_err = "b200enum_run: solution buffer overflow";
    return B200ENUM_EOVERFLOW;
  }
  return B200ENUM_OK;
}

}  // extern "C"
