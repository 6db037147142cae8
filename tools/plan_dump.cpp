// Development aid / unit-test helper: prints the Gao-Mateer constants summary and the pass
// schedule that bch_fft_b200/csrc/gm_plan.h produces for a field degree.
//   g++ -O2 -std=c++17 -I bch_fft_b200/csrc tools/plan_dump.cpp -o /tmp/plan_dump && /tmp/plan_dump 16 16 11
#include <cstdio>
#include <cstdlib>
#include "gm_plan.h"
using namespace bchfft;

static const uint32_t kPoly[26] = {0, 0, 0x7, 0xB, 0x13, 0x25, 0x43, 0x83, 0x171, 0x211, 0x409, 0x805, 0x1099,
                                   0x201B, 0x5803, 0x8003, 0x1002D, 0x20009, 0x40081, 0x80063, 0x100009,
                                   0x200005, 0x400003, 0x800021, 0x100001B, 0x2000009};

int main(int argc, char **argv)
{
  const int m = argc > 1 ? atoi(argv[1]) : 16;
  const int K = argc > 2 ? atoi(argv[2]) : m;
  const int t = argc > 3 ? atoi(argv[3]) : 11;
  const uint32_t n = 1u << m;
  std::vector<uint32_t> ex(n), lg(n);
  uint32_t v = 1;
  for (uint32_t i = 0; i + 1 < n; i++) {
    ex[i] = v;
    lg[v] = i;
    v <<= 1;
    if (v >> m) v ^= kPoly[m];
  }
  ex[n - 1] = 0;
  lg[0] = n - 1;
  GmConstants c = gm_constants(m, ex.data(), lg.data());
  int chain = 0;
  for (int d = 0; d < m; d++)
    if ((c.twist_free >> d) & 1u) chain++;
  // sanity: the basis spans the field and the position tables are inverse permutations
  bool perm_ok = true;
  for (uint32_t k = 0; k < n; k++) perm_ok &= c.elem_of_pos[c.pos_of_elem[k]] == k;
  GmPlan plan = gm_plan(m, K, t, c.twist_free, !getenv("BCHFFT_NO_CANTOR_PREFIX"));
  int ntw = 0, nty = 0, nbf = 0, ntx = 0;
  for (auto &p : plan.passes)
    for (uint32_t i = 0; i < p.nops; i++) {
      ntw += p.ops[i].type == OP_TW;
      nty += p.ops[i].type == OP_TY;
      nbf += p.ops[i].type == OP_BF;
      ntx += p.ops[i].type == OP_TX;
    }
  printf("m=%d K=%d t=%d twist_free_depths=%d perm_ok=%d passes=%zu tw_ops=%d ty_ops=%d bf_ops=%d tx_ops=%d\n", m, K, t,
         chain, (int)perm_ok, plan.passes.size(), ntw, nty, nbf, ntx);
  for (auto &p : plan.passes) {
    uint32_t w = (((1u << p.n0) - 1u) << p.a0) | (p.n1 ? (((1u << p.n1) - 1u) << p.a1) : 0u);
    printf("  pass dom=%u t=%u window=0x%x ops=%u backward=%u\n", p.dom_bits, p.t, w, p.nops, p.backward);
  }
  GmPlanPM pm;
  const int t1 = argc > 4 ? atoi(argv[4]) : 14;
  if (K == m && gm_plan_pm(m, t1, t, c.twist_free, pm)) {
    auto dump = [](const char *name, const std::vector<PassDesc> &v) {
      for (auto &p : v) {
        uint32_t w = (((1u << p.n0) - 1u) << p.a0) | (p.n1 ? (((1u << p.n1) - 1u) << p.a1) : 0u);
        printf("  pm %s pass t=%u window=0x%x ops:", name, p.t, w);
        for (uint32_t i = 0; i < p.nops; i++)
          printf(" %s(%u,%u)", p.ops[i].type == OP_TX ? "TX" : p.ops[i].type == OP_BF ? "BF" : "??",
                 p.ops[i].type == OP_TX ? p.ops[i].lo : p.ops[i].d, p.ops[i].type == OP_TX ? p.ops[i].d : 0u);
        printf("\n");
      }
    };
    dump("line", pm.line_passes);
    dump("plane", pm.plane_passes);
    dump("bf", pm.bf_passes);
  }
  return perm_ok ? 0 : 1;
}
