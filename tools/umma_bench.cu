// umma_bench.cu -- microbenchmark of tcgen05.mma.kind::i8 issue cost on sm_100a, to size the sweep
// kernel's tile shape.  One CTA per SM, one issuing thread, operands resident in
// SMEM (no TMA), SWIZZLE_128B K-major tiles walked exactly like k_sweep_tc does (4 UMMAs of K=32 bytes
// per 128-byte row, `boxes` tiles in a ring).  Optional "hammer" warps read the same tiles with
// conflict-free LDS.32 to emulate the burden-collapse warps competing for the SMEM port.
//
//   umma_bench M N [a_mode] [hammer_warps] [same_tile] [iters] [nacc] [layout]
//     layout: 0 = SWIZZLE_128B rows (4 K-slices per row), 1 = SWIZZLE_32B (one 32-byte row per K-slice tile),
//             2 = no swizzle (8x16-byte core matrices)
//     nacc: number of TMEM accumulators the UMMAs rotate over (1 = every UMMA accumulates into the same D)
//     a_mode: 0 = A from SMEM (SS), 1 = A from TMEM (TS)
//     same_tile: 1 = B descriptor starts at the A tile (B = [A ; extra rows]) as in the sweep
// prints: clk per UMMA (K=32), clk per 128-sample box, LDS bytes/clk achieved by the hammer warps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint64_t desc_any(uint32_t smem_addr, int layout) {
  if (layout == 0) return desc_sw128(smem_addr);
  if (layout == 1) {   // SWIZZLE_32B: 8 rows x 32 B atoms, SBO = 256 B
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = 16u | (1u << 14) | (6u << 29);
    return ((uint64_t)hi << 32) | lo;
  }
  // no swizzle: core matrix = 8 rows x 16 B contiguous (128 B); K-adjacent cores LBO = 128 B apart,
  // 8-row groups SBO = 256 B apart
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (8u << 16);
  const uint32_t hi = 16u | (1u << 14) | (0u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\nmov.s32 %0, %%rx;\n}\n"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

constexpr int kBoxes = 4;
constexpr int kRowsMax = 256 + 128;   // A tile then B tile (or shared)
constexpr int kBoxBytes = kRowsMax * 128;

__global__ void __launch_bounds__(64 + 32 * 8, 1)
k_bench(int M, int N, int a_mode, int hammer, int same_tile, int iters, int nacc, int layout, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kBoxes * kBoxBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint32_t st = smem_u32(tiles);
      const uint32_t boff = same_tile ? 0u : (uint32_t)(128 * 128);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t box = st + (uint32_t)(it % kBoxes) * kBoxBytes;
        const uint64_t da = desc_any(box, layout), db = desc_any(box + boff, layout);
        // K-slice k: +32 B inside the 128-byte row (SW128) or the next rows x 32 B tile (SW32 / none)
        const uint64_t kstep = layout == 0 ? 2u : (uint64_t)((kRowsMax * 32) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem_base + (uint32_t)(((it * 4 + k) % nacc) * N);
          if (a_mode == 0)
            umma_ss(d, da + kstep * k, db + kstep * k, idesc, (it >= 1) ? 1u : 0u);
          else
            umma_ts(d, tmem_base + 480u + 8u * k, db + kstep * k, idesc, (it >= 1) ? 1u : 0u);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      t1 = clock64();
      stop = 1;
      out[blockIdx.x * 4 + 0] = t1 - t0;
    }
  } else if (warp >= 2 && warp < 2 + hammer) {
    // conflict-free LDS.32 over the tiles until the issuer is done
    uint32_t acc = 0;
    long long n = 0;
    const long long h0 = clock64();
    while (!stop) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<volatile uint32_t*>(tiles + ((warp * 7 + r) % 32) * 1024 + j * 128 + lane * 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += w[j];
      }
      n += 64;
    }
    const long long h1 = clock64();
    if (lane == 0) {
      out[blockIdx.x * 4 + 1] = acc;   // keep the loads alive
      atomicAdd((unsigned long long*)&out[blockIdx.x * 4 + 2], (unsigned long long)(n * 128));
      out[blockIdx.x * 4 + 3] = h1 - h0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512));
  }
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 80;
  const int a_mode = argc > 3 ? atoi(argv[3]) : 0, hammer = argc > 4 ? atoi(argv[4]) : 0;
  const int same = argc > 5 ? atoi(argv[5]) : 1, iters = argc > 6 ? atoi(argv[6]) : 4000;
  const int nacc = argc > 7 ? atoi(argv[7]) : 1;
  const int layout = argc > 8 ? atoi(argv[8]) : 0;
  const int grid = 148;
  long long* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(long long) * 4 * grid);
  const int smem = kBoxes * kBoxBytes + 1024;
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_out, 0, sizeof(long long) * 4 * grid);
    k_bench<<<grid, 64 + 32 * 8, smem>>>(M, N, a_mode, hammer, same, iters, nacc, layout, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("M=%d N=%d a_mode=%d hammer=%d: CUDA error %s\n", M, N, a_mode, hammer, cudaGetErrorString(e));
      return 1;
    }
  }
  long long h[4 * 148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double clk = 0, lds = 0;
  for (int b = 0; b < grid; ++b) {
    clk += (double)h[b * 4];
    if (hammer) lds += (double)h[b * 4 + 2] / (double)h[b * 4 + 3];
  }
  clk /= grid;
  printf("M=%3d N=%3d a=%s hammer=%d same_tile=%d nacc=%d layout=%d : %7.1f clk/UMMA(K=32)  %7.1f clk/box(128 samples)  LDS %.1f B/clk/SM\n", M, N,
         a_mode ? "TMEM" : "SMEM", hammer, same, nacc, layout, clk / (4.0 * iters), clk / iters, hammer ? lds / grid : 0.0);
  return 0;
}
