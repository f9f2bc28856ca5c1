/* enum_oracle.c — CPU ORACLE for the BKZ enumeration path, TEST INFRASTRUCTURE ONLY (see gso_oracle.c header).
 *
 * Restates the reference's Schnorr-Euchner enumeration for the case BKZ uses (SVP, primal, no sub-solutions, no CVP
 * reset): EnumerationDyn::prepare_enumeration (enum/enumerate.cpp:161-216), EnumerationBase::enumerate_loop
 * (enum/enumerate_base.cpp:152-254), next_pos_up (enum/enumerate_base.h:145-171), set_bounds / process_solution
 * (enum/enumerate.cpp:218-239) with the FastEvaluator default strategy "keep the best 1" (enum/evaluator.h:122-156).
 * Inputs are exactly what the external-enumerator hook receives (enum/enumerate_ext.cpp:91-148): mut[i*d+j] = mu(j,i)
 * for j > i, rdiag and maxdist already normalised by 2^-normexp.
 *
 * Centres are recomputed as the chain  c_k = ((0 - x[d-1]*mut[k][d-1]) - x[d-2]*mut[k][d-2]) - ... - x[k+1]*mut[k][k+1]
 * which is bit-identical to the reference's cached center_partsums (the cache memoises prefixes of this same chain,
 * enumerate_base.cpp:53-62,232-241), so pruning decisions and node counts match the reference's internal enumerator.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OENUM_MAXD 256

long oenum_svp_ex(int d, const double *mut_in, const double *rdiag_in, const double *pruning, double maxdist, int shrink,
                  int dual, int findsubsols, double *sol, double *best, uint64_t *nodes, double *subdist, double *subsol);

/* shrink != 0: BEST_1 evaluator (maxdist := dist of each new solution, as BKZ's FastEvaluator does);
 * shrink == 0: fixed radius (counts every leaf inside the bound; *nsols = number of leaves).
 * Returns the number of solutions reported; best solution in sol[], its dist in *best. nodes[d] per level. */
long oenum_svp(int d, const double *mut, const double *rdiag, const double *pruning, double maxdist, int shrink,
               double *sol, double *best, uint64_t *nodes)
{
  return oenum_svp_ex(d, mut, rdiag, pruning, maxdist, shrink, 0, 0, sol, best, nodes, NULL, NULL);
}

/* The general form of the hook (enumerate_ext_api.h:88-92): dual != 0 enumerates the dual of the block — mut_in /
 * rdiag_in are the PRIMAL block's, reversed and inverted here as EnumerationDyn::enumerate does (enumerate.cpp:100-113),
 * the centre chains run over alpha_j = x_j - c_j (enumerate_base.cpp:64-68,232-236) and the solution is handed back in
 * the block's own order (enumerate.cpp:150-154).  findsubsols != 0 keeps, per level k, the shortest partial vector
 * with 0 < length < rdiag[k] (enumerate_base.cpp:36-40, subsoldists = rdiag, enumerate.cpp:141): subdist[k] (-1 = none)
 * and subsol[k*d + j] (zero for j < k; enumeration-order indices, as process_subsolution stores them). */
long oenum_svp_ex(int d, const double *mut_in, const double *rdiag_in, const double *pruning, double maxdist, int shrink,
                  int dual, int findsubsols, double *sol, double *best, uint64_t *nodes, double *subdist, double *subsol)
{
  double x[OENUM_MAXD + 1], center[OENUM_MAXD + 1], partdist[OENUM_MAXD + 1], bounds[OENUM_MAXD];
  double alpha[OENUM_MAXD + 1], ssd[OENUM_MAXD];
  double *mut_t = NULL, *rd_t = NULL;
  const double *mut = mut_in, *rdiag = rdiag_in;
  if (dual)
  {
    mut_t = (double *)calloc((size_t)d * d, sizeof(double));
    rd_t  = (double *)calloc(d, sizeof(double));
    for (int i = 0; i < d; i++)
      rd_t[d - 1 - i] = 1.0 / rdiag_in[i];
    for (int i = 0; i < d; i++)
      for (int j = i + 1; j < d; j++)
        mut_t[(size_t)(d - 1 - j) * d + (d - 1 - i)] = -mut_in[(size_t)i * d + j];
    mut = mut_t, rdiag = rd_t;
  }
  for (int i = 0; i < d; i++)
  {
    ssd[i]   = rdiag[i];
    alpha[i] = 0.0;
    if (findsubsols && subdist)
      subdist[i] = -1.0;
  }
#define CHAINV(j) (dual ? alpha[j] : x[j])
  int dx[OENUM_MAXD + 1], ddx[OENUM_MAXD + 1];
  long nsols = 0;
  int k, k_end = d;
  *best = -1.0;
  memset(nodes, 0, sizeof(uint64_t) * d);
  for (int i = 0; i < d; i++)
    bounds[i] = (pruning ? pruning[i] : 1.0) * maxdist; /* set_bounds */

  /* prepare_enumeration, SVP without subtree: all centres are 0 so x = 0, partdist = 0 all the way down */
  {
    double newdist = 0.0;
    for (k = d - 1; k >= 0 && newdist <= maxdist; --k)
    {
      double nc = 0.0;
      for (int j = d - 1; j > k; --j)
        nc = nc - CHAINV(j) * mut[(size_t)k * d + j];
      x[k]        = round(nc);
      center[k]   = nc;
      partdist[k] = newdist;
      dx[k] = ddx[k] = (nc >= x[k]) ? 1 : -1;
      double a       = x[k] - nc;
      alpha[k]       = a; /* prepare_enumeration, enumerate.cpp:206-210 */
      newdist        = newdist + a * a * rdiag[k];
    }
    x[0] = 1; /* excludes the zero vector */
    ++k;
  }
  if (k >= k_end)
  {
    free(mut_t), free(rd_t);
    return 0;
  }
  partdist[k_end] = 0.0;
  for (int i = k + 1; i < k_end; i++)
    nodes[i]--; /* node-count compensation of the initial descent, enumerate_base.cpp:165-183 */
  k = k_end - 1;

  int finished = 0;
  while (!finished)
  {
    double alphak  = x[k] - center[k];
    double newdist = partdist[k] + alphak * alphak * rdiag[k];
    int up         = 0;
    if (newdist <= bounds[k])
    {
      ++nodes[k];
      alpha[k] = alphak;
      if (findsubsols && newdist < ssd[k] && newdist != 0.0)
      {
        ssd[k] = newdist;
        if (subdist)
        {
          subdist[k] = newdist;
          for (int j = 0; j < d; j++)
            subsol[(size_t)k * d + j] = j < k ? 0.0 : x[j];
        }
      }
      --k;
      if (k < 0)
      {
        if (newdist > 0.0)
        {
          nsols++;
          if (*best < 0 || newdist < *best)
          {
            *best = newdist;
            for (int j = 0; j < d; j++)
              sol[dual ? d - 1 - j : j] = x[j];
          }
          if (shrink)
          {
            maxdist = newdist; /* BEST_1: new radius = this solution's length */
            for (int i = 0; i < d; i++)
              bounds[i] = (pruning ? pruning[i] : 1.0) * maxdist;
          }
        }
        up = 1;
      }
      else
      {
        double nc = 0.0;
        for (int j = d - 1; j > k; --j)
          nc = nc - CHAINV(j) * mut[(size_t)k * d + j];
        center[k]   = nc;
        partdist[k] = newdist;
        x[k]        = round(nc);
        dx[k] = ddx[k] = (nc >= x[k]) ? 1 : -1;
      }
    }
    else
      up = 1;
    if (up)
    {
      /* next_pos_up */
      ++k;
      if (partdist[k] != 0.0)
      {
        x[k] += dx[k];
        ddx[k] = -ddx[k];
        dx[k]  = ddx[k] - dx[k];
      }
      else
      {
        if (k >= k_end)
          finished = 1;
        else
          ++x[k]; /* SVP: break the +/- symmetry at the top non-zero coefficient */
      }
    }
  }
#undef CHAINV
  free(mut_t), free(rd_t);
  return nsols;
}
