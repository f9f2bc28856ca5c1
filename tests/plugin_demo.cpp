// plugin_demo — TEST PROGRAM: the UNMODIFIED reference library (oracle/_ref/libfplll.so) running its own
// bkz_reduction with the device enumerator installed through fplll's plugin hook (set_external_enumerator,
// fplll/enum/enumerate_ext.h:100) by fplll_b200/csrc/fplll_extenum_adapter.cpp.  This is the drop-in scenario of
// INTEGRATION.md, end to end.  usage: plugin_demo IN.txt OUT.txt BLOCK FLAGS MAXLOOPS default|none b200|enumlib
#include <fplll/fplll.h>
#include <chrono>
#include <fstream>
#include <iostream>
extern "C" void b200_enum_register(int ngpus);
using namespace fplll;
int main(int argc, char **argv)
{
  if (argc < 8)
    return 2;
  ZZ_mat<mpz_t> B;
  std::ifstream f(argv[1]);
  f >> B;
  const int bs = atoi(argv[3]), fl = atoi(argv[4]), ml = atoi(argv[5]);
  std::vector<Strategy> strategies;
  if (std::string(argv[6]) == "default")
    strategies = load_strategies_json(strategy_full_path("default.json"));
  if (std::string(argv[7]) == "b200")
    b200_enum_register(1);
  BKZParam param(bs, strategies);
  param.flags     = fl;
  param.max_loops = ml;
  auto t0         = std::chrono::steady_clock::now();
  int st          = bkz_reduction(&B, NULL, param, FT_DOUBLE, 0);
  double sec      = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::ofstream o(argv[2]);
  o << B << std::endl;
  printf("plugin_demo enumerator=%s status=%d sec=%.6f\n", argv[7], st, sec);
  return 0;
}
