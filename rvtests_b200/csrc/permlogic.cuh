// permlogic.cuh -- the host-testable core of the permutation test (perm.cuh): glibc's rand() as a linear
// recurrence with polynomial jump-ahead, and Fisher-Yates resolved as chains.  __host__ __device__, no CUDA
// headers: tests/hostcheck compiles it with g++ and pins it against glibc's own rand() and a literal shuffle.
#pragma once
#include <stdint.h>

#include "mathdev.cuh"

#ifndef __CUDACC__
#ifndef __restrict__
#define __restrict__
#endif
#endif

namespace rvt {

// ---- glibc rand() as linear algebra over Z/2^32 ------------------------------------------------
constexpr int kLfgDeg = 31;
constexpr int kLfgWarm = 341;   // y-index of the first output: r[344] of the textbook form, y_m = r[m+3]
struct LfgPoly {
  uint32_t c[kLfgDeg];   // residue modulo z^31 - z^28 - 1
};

RVT_HDN LfgPoly lfg_mul(const LfgPoly& a, const LfgPoly& b) {
  uint32_t t[2 * kLfgDeg - 1];
  for (int i = 0; i < 2 * kLfgDeg - 1; ++i) t[i] = 0;
  for (int i = 0; i < kLfgDeg; ++i)
    for (int j = 0; j < kLfgDeg; ++j) t[i + j] += a.c[i] * b.c[j];
  for (int d = 2 * kLfgDeg - 2; d >= kLfgDeg; --d) {   // z^d = z^(d-3) + z^(d-31)
    t[d - 3] += t[d];
    t[d - kLfgDeg] += t[d];
  }
  LfgPoly r;
  for (int i = 0; i < kLfgDeg; ++i) r.c[i] = t[i];
  return r;
}
RVT_HDN LfgPoly lfg_one() {
  LfgPoly r;
  for (int i = 0; i < kLfgDeg; ++i) r.c[i] = 0;
  r.c[0] = 1;
  return r;
}
RVT_HDN LfgPoly lfg_pow(uint64_t n) {   // z^n
  LfgPoly base = lfg_one(), acc = lfg_one();
  base.c[0] = 0;
  base.c[1] = 1;
  while (n) {
    if (n & 1) acc = lfg_mul(acc, base);
    n >>= 1;
    if (n) base = lfg_mul(base, base);
  }
  return acc;
}
// first 61 values y_0..y_60 after srand(seed) (glibc stdlib/random_r.c, TYPE_3): r_0 = seed,
// r_i = 16807 r_{i-1} mod (2^31 - 1) for i < 31, r_{31..33} = r_{0..2}, then the recurrence; y_m = r_{m+3}
RVT_HDN void lfg_seed_window(uint32_t seed, uint32_t* y /*[2*kLfgDeg-1]*/) {
  uint32_t r[3 + 2 * kLfgDeg - 1];
  if (seed == 0) seed = 1;
  r[0] = seed;
  for (int i = 1; i < 31; ++i) {
    long long w = (16807LL * (int32_t)r[i - 1]) % 2147483647LL;
    if (w < 0) w += 2147483647LL;
    r[i] = (uint32_t)w;
  }
  for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
  for (int i = 34; i < 3 + 2 * kLfgDeg - 1; ++i) r[i] = r[i - 31] + r[i - 3];
  for (int m = 0; m < 2 * kLfgDeg - 1; ++m) y[m] = r[m + 3];
}
// window y_{n..n+60} from the poly z^n and the seed window
RVT_HDN void lfg_window_at(const LfgPoly& zn, const uint32_t* y0 /*[61]*/, uint32_t* w /*[61]*/) {
  for (int k = 0; k < kLfgDeg; ++k) {
    uint32_t s = 0;
    for (int j = 0; j < kLfgDeg; ++j) s += zn.c[j] * y0[j + k];
    w[k] = s;
  }
  for (int k = kLfgDeg; k < 2 * kLfgDeg - 1; ++k) w[k] = w[k - 31] + w[k - 3];
}

// ---- Fisher-Yates as chains --------------------------------------------------------------------
constexpr uint32_t kFyNil = 0xFFFFFFFFu;
// steps that target position q form the list head[q] -> link[.] -> ...; the smallest member > t, or kFyNil
RVT_HD uint32_t fy_succ(const uint32_t* __restrict__ head, const uint32_t* __restrict__ link, uint32_t q, uint32_t t) {
  uint32_t s = kFyNil;
  for (uint32_t e = head[q]; e != kFyNil; e = link[e])
    if (e > t && e < s) s = e;
  return s;
}
// index (into the vector BEFORE this shuffle) of the value that ends at position i; j = j_i (ignored for i == 0)
RVT_HD uint32_t fy_root(const uint32_t* __restrict__ head, const uint32_t* __restrict__ link, uint32_t i, uint32_t j) {
  uint32_t q = (i == 0) ? 0u : j, t = i;
  for (;;) {
    const uint32_t s = fy_succ(head, link, q, t);
    if (s == kFyNil) return q;
    q = s;
    t = s;
  }
}

}  // namespace rvt
