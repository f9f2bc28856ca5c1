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

/* shrink != 0: BEST_1 evaluator (maxdist := dist of each new solution, as BKZ's FastEvaluator does);
 * shrink == 0: fixed radius (counts every leaf inside the bound; *nsols = number of leaves).
 * Returns the number of solutions reported; best solution in sol[], its dist in *best. nodes[d] per level. */
long oenum_svp(int d, const double *mut, const double *rdiag, const double *pruning, double maxdist, int shrink,
               double *sol, double *best, uint64_t *nodes)
{
  double x[OENUM_MAXD + 1], center[OENUM_MAXD + 1], partdist[OENUM_MAXD + 1], bounds[OENUM_MAXD];
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
        nc = nc - x[j] * mut[(size_t)k * d + j];
      x[k]        = round(nc);
      center[k]   = nc;
      partdist[k] = newdist;
      dx[k] = ddx[k] = (nc >= x[k]) ? 1 : -1;
      double a       = x[k] - nc;
      newdist        = newdist + a * a * rdiag[k];
    }
    x[0] = 1; /* excludes the zero vector */
    ++k;
  }
  if (k >= k_end)
    return 0;
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
              sol[j] = x[j];
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
          nc = nc - x[j] * mut[(size_t)k * d + j];
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
  return nsols;
}
