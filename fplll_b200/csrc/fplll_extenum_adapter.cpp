// fplll_extenum_adapter.cpp — the reference-side binding of the device enumerator.
//
// Compiled AGAINST fplll's own headers (it needs std::function typedefs from fplll/enum/enumerate_ext_api.h) and
// linked into the application that links libfplll; everything below the typedef'd signature is the plain C-ABI of
// include/b200enum.h.  Registration is one call: b200_enum_register(ngpus) -> fplll::set_external_enumerator(...)
// (fplll/enum/enumerate_ext.h:100), or at configure time  --with-extenum-func=b200_enumerate  (configure.ac:187-210).
//
// Not part of libb200enum.so: /root/reference's headers do not exist on the GPU box, so this file is built only
// where fplll is installed (tests/ builds it against the reference build of the development container to prove it links and
// runs — see INTEGRATION.md).
#include <fplll/fplll.h>

#include <array>
#include <stdexcept>
#include <vector>

#include "../../include/b200enum.h"

namespace {

int g_ngpus = 1;

struct Callbacks
{
  std::function<fplll::extenum_cb_process_sol> *sol;
  std::function<fplll::extenum_cb_process_subsol> *subsol;
};

double trampoline(void *ctx, double dist, const double *sol)
{
  auto *c = static_cast<Callbacks *>(ctx);
  return (*c->sol)(dist, const_cast<double *>(sol));  // extenum_cb_process_sol returns the new bound (enumerate_ext_api.h:62)
}

void sub_trampoline(void *ctx, double dist, const double *subsol, int offset)
{
  auto *c = static_cast<Callbacks *>(ctx);
  (*c->subsol)(dist, const_cast<double *>(subsol), offset);  // enumerate_ext_api.h:70-71
}

}  // namespace

// exactly fplll::extenum_fc_enumerate (enumerate_ext_api.h:88-92)
std::array<uint64_t, FPLLL_EXTENUM_MAX_EXTENUM_DIM>
b200_enumerate(const int dim, fplll::enumf maxdist, std::function<fplll::extenum_cb_set_config> cbfunc,
               std::function<fplll::extenum_cb_process_sol> cbsol,
               std::function<fplll::extenum_cb_process_subsol> cbsubsol, bool dual, bool findsubsols)
{
  std::array<uint64_t, FPLLL_EXTENUM_MAX_EXTENUM_DIM> ret{};
  if (dim < 2 || dim > B200ENUM_MAX_DIM)
  {
    ret[0] = ~uint64_t(0);  // "not supported": fplll uses its own enumerator (enumerate_ext.cpp:88); dual and
    return ret;             // sub-solution requests ARE served (enumlib declines dual, enumlib.cpp:98-104)
  }
  std::vector<double> mut((size_t)dim * dim, 0.0), rdiag(dim), pruning(dim);
  cbfunc(mut.data(), dim, /*mutranspose=*/true, rdiag.data(), pruning.data());
  std::vector<int> devs(g_ngpus);
  for (int i = 0; i < g_ngpus; i++)
    devs[i] = i;
  std::vector<uint64_t> nodes(dim, 0);
  Callbacks cbs{&cbsol, &cbsubsol};
  const int flags = (dual ? B200ENUM_DUAL : 0) | (findsubsols ? B200ENUM_FINDSUBSOLS : 0);
  const int rc    = b200enum_run_ex(dim, maxdist, mut.data(), rdiag.data(), pruning.data(), flags, devs.data(), g_ngpus,
                                    0, 1, trampoline, findsubsols ? sub_trampoline : nullptr, &cbs, nodes.data(), nullptr);
  if (rc != B200ENUM_OK)  // a broken GPU must not silently turn into a CPU enumeration
    throw std::runtime_error(std::string("b200_enumerate: ") + b200enum_last_error());
  for (int i = 0; i < dim; i++)
    ret[i] = nodes[i];
  return ret;
}

extern "C" void b200_enum_register(int ngpus)
{
  g_ngpus = ngpus > 0 ? ngpus : 1;
  fplll::set_external_enumerator(b200_enumerate);
}
