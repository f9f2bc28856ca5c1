// stream_table_check — TEST PROGRAM (host only): prints the chunk table the TMA streaming update kernel builds
// (fplll_b200/csrc/gso_stream.cuh: stream_shape / stream_num_chunks / stream_desc) for one (d, n, i, last_j), one line
// per chunk: "e map bytes c1 c2".  tests/test_stream_table_cpu.py checks it against an independent walk of the consumer's
// loop nest and against the algorithmic bytes of update_gso_row.
// build: nvcc -std=c++17 -o stream_table_check tests/stream_table_check.cu     usage: stream_table_check d n i last_j
#include "../fplll_b200/csrc/gso_stream.cuh"
#include <cstdio>
#include <cstdlib>
using namespace b200;
int main(int argc, char **argv)
{
  if (argc < 5)
    return 2;
  const int d = atoi(argv[1]), n = atoi(argv[2]), i = atoi(argv[3]), last_j = atoi(argv[4]);
  const StreamShape sh = stream_shape(i, last_j);
  const int NC         = stream_num_chunks(sh, n);
  printf("shape d=%d jl=%d P=%d rows_last=%d NC=%d cols=%d stages=%d\n", d, sh.jl, sh.P, sh.rows_last, NC, ST_COLS, ST_STAGES);
  for (int e = 0; e < NC; e++)
  {
    const StreamDesc q = stream_desc(sh, n, ld_b(n), e);
    printf("%d %d %d %d %d\n", e, q.map, q.bytes, q.c1, q.c2);
  }
  return 0;
}
