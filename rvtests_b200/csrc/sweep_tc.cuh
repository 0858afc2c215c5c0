// sweep_tc.cuh -- K1 on the 5th-generation tensor cores (placeholder until the kernel lands).
#pragma once
#include "common.cuh"
namespace rvt {
struct TcSegments {
  void* encode = nullptr;
  char why[128] = "tensor-core sweep not built yet";
};
inline int tc_init(TcSegments*, char*, size_t) { return 0; }
inline int tc_bind_null(TcSegments*, const int8_t*, int, int64_t, int64_t, char*, size_t) { return 0; }
inline int tc_bind_segment(TcSegments*, int, const int8_t*, int64_t, int64_t, int64_t, char*, size_t) { return 0; }
inline bool tc_usable(const TcSegments*, const GeneDesc*, int) { return false; }
inline int tc_launch(TcSegments*, const GeneDesc*, const GeneDesc*, int, const uint8_t*, const NullModel*, int64_t, int,
                     int, int64_t, SweepPartial*, unsigned int*, int, cudaStream_t, char*, size_t) {
  return -4;
}
}  // namespace rvt
