// Micro-benchmark of tcgen05.mma issue/execution cost for the instruction shapes of the fused InfoNCE kernel
// (csrc/infonce_tc.cu): cycles per MMA for back-to-back MMAs of one shape, and for the S/O interleaving of the tile loop.
// Build (here, cross-compiled):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I robust-multimodal-contrastive-learning_b200/csrc \
//                                     tools/mma_bench.cu -o build/mma_bench -lcuda
// Run on the GPU box:            build/mma_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

namespace rmcl { void set_error(const char*, ...) {} int sm_count() { return 148; } }
using namespace rmcl::tcx;

struct Case {
  int kind;      // 0: TS (A in TMEM), 1: SS
  int n;         // MMA N
  int b_mn;      // B MN-major (S GEMM) or K-major (O GEMM)
  int per_group; // MMAs per commit group
  int groups;
  int interleave;  // 1: alternate [16 x TS N64 MN-major] and [4 x SS N256 K-major] like the tile loop
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(Case c, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  // deterministic non-NaN contents (bf16 0x3c00-ish small numbers)
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    const uint32_t sbase = smem_u32(sm);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms the instruction cache
      t0 = clock64();
      if (!c.interleave) {
        const uint32_t idesc = make_idesc(128, c.n, c.b_mn);
        for (int g = 0; g < c.groups; ++g) {
          if (elect_one()) {
#pragma unroll 1
            for (int s = 0; s < c.per_group; ++s) {
              const uint32_t off = (uint32_t)(s & 15) * 2048u;
              const uint64_t bd = c.b_mn ? make_sw128_desc(sbase + 65536 + off, 32768, 1024)
                                         : make_sw128_desc(sbase + 65536 + (s & 3) * 32 + (uint32_t)((s >> 2) & 3) * 32768u, 16, 1024);
              if (c.kind == 0) tc_mma_ts(tmem + 256 + (g & 1) * 0, tmem + (s & 15) * 8, bd, idesc, 1);
              else {
                const uint64_t ad = make_sw128_desc(sbase + (s & 3) * 32, 16, 1024);
                tc_mma_ss(tmem + 256, ad, bd, idesc, 1);
              }
            }
          }
          __syncwarp();
        }
      } else {
        const uint32_t idS = make_idesc(128, 64, 1), idO = make_idesc(128, 256, 0);
        for (int g = 0; g < c.groups; ++g) {
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < 16; ++s) {
              const uint64_t bd = make_sw128_desc(sbase + 65536 + s * 2048, 32768, 1024);
              tc_mma_ts(tmem + 384 + (g & 1) * 64, tmem + s * 8, bd, idS, s > 0);
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const uint64_t ad = make_sw128_desc(sbase + (g & 1) * 16384 + s * 32, 16, 1024);
              const uint64_t bd = make_sw128_desc(sbase + 65536 + 32768 + s * 32, 16, 1024);
              tc_mma_ss(tmem + 128, ad, bd, idO, 1);
            }
          }
          __syncwarp();
        }
      }
      t1 = clock64();
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t2 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;   // issue
      out[1] = t2 - t0;   // issue + drain
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The tile loop's MMA stream (16 x TS N64 + 4 x SS N256 per tile, as above) with the OTHER traffic of the fused kernel running
// beside it at the rate of one burst per `period` cycles:
//   bit 0  four warps (one per TMEM lane quadrant) read 64 accumulator columns with tcgen05.ld        (softmax reading S)
//   bit 1  four warps store one 128-byte row each per burst into a swizzled P-like tile (16 KB)          (softmax writing P)
//   bit 2  one lane streams 32 KB per burst from global into the ring with cp.async.bulk                (the TMA producer)
//   bit 3  the tcgen05.ld warps also run 64 ex2 + cvt per row per burst                                  (the exponentials)
//   bit 4  a tcgen05.commit behind every S GEMM and two behind every O GEMM (to barriers nobody waits on)  (the kernel's protocol)
//   bit 5  ... and the issuer waits for the commit of the S GEMM two tiles back before each S GEMM          (s_full-like dependency)
__global__ void __launch_bounds__(448, 1) mma_stress_kernel(int mode, int period, int tiles, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar, lbar, dummy[4], sbar[2], done3[3];
  __shared__ volatile int goflag;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (mode & 64) {   // random bf16 pairs in (-2, 2): operands that toggle like real data
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x807f807fu) | 0x3f003f00u | ((h >> 3) & 0x00800080u);
    }
    reinterpret_cast<uint32_t*>(sm)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&lbar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1);
    mbar_init(&sbar[0], 1);
    mbar_init(&sbar[1], 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&done3[i], 1);
      mbar_arrive(&done3[i]);     // phase 0 complete: a parity-0 wait succeeds at once
    }
    goflag = 1;
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t sbase = smem_u32(sm);
  if ((mode & 64) && warp >= 4 && warp < 8) {   // random Q^ (A operand of the S GEMMs) in tensor memory columns 0..127
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint32_t h = (uint32_t)(threadIdx.x * 131 + ch * 32 + j) * 2654435761u + blockIdx.x * 977u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        w[j] = (h & 0x807f807fu) | 0x3d003d00u;
      }
      tc_st32(tl + ch * 32, w);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idS = make_idesc(128, 64, 1), idO = make_idesc(128, 256, 0);
    long long t0 = 0, t2 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int g = 0; g < tiles; ++g) {
        const int gg = rep * tiles + g;
        if ((mode & 32) && gg >= 2) mbar_wait(&sbar[gg & 1], ((gg - 2) >> 1) & 1);
        if (mode & 128) {          // the kernel's issuer between two bursts: probe three (complete) barriers, fence
          while ((mbar_try_wait3(&done3[0], 0, &done3[1], 0, &done3[2], 0) & 7u) != 7u) {}
          tc_fence_after();
        }
        if (mode & 256) {          // a flag posted by a helper warp instead of the probe
          int f;
          do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(f) : "r"(smem_u32((const void*)&goflag)) : "memory"); } while (f == 0);
          tc_fence_after();
        }
        if (mode & 512) tc_fence_after();
        if (mode & 1024) {         // single elect block per tile (merged burst), as the kernel issues it now
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < 16; ++s) {
              const uint64_t bd = make_sw128_desc(sbase + 65536 + s * 2048, 32768, 1024);
              tc_mma_ts(tmem + 384 + (g & 1) * 64, tmem + s * 8, bd, idS, s > 0);
            }
            tc_commit(&dummy[0]);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const uint64_t ad = make_sw128_desc(sbase + (g & 1) * 16384 + s * 32, 16, 1024);
              const uint64_t bd = make_sw128_desc(sbase + 65536 + 32768 + s * 32, 16, 1024);
              tc_mma_ss(tmem + 128, ad, bd, idO, 1);
            }
            tc_commit(&dummy[1]);
            tc_commit(&dummy[2]);
          }
          __syncwarp();
          continue;
        }
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 16; ++s) {
            const uint64_t bd = make_sw128_desc(sbase + 65536 + s * 2048, 32768, 1024);
            tc_mma_ts(tmem + 384 + (g & 1) * 64, tmem + s * 8, bd, idS, s > 0);
          }
          if (mode & 32) tc_commit(&sbar[gg & 1]);
          else if (mode & 16) tc_commit(&dummy[0]);
        }
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint64_t ad = make_sw128_desc(sbase + (g & 1) * 16384 + s * 32, 16, 1024);
            const uint64_t bd = make_sw128_desc(sbase + 65536 + 32768 + s * 32, 16, 1024);
            tc_mma_ss(tmem + 128, ad, bd, idO, 1);
          }
          if (mode & 48) {
            tc_commit(&dummy[1]);
            tc_commit(&dummy[2]);
          }
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t2 = clock64();
    }
    if (lane == 0) {
      stop = 1;
      if (blockIdx.x == 0) out[0] = t2 - t0;
    }
  } else if (warp >= 4 && warp < 8) {
    if (mode & 9) {
      const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384;
      long long next = clock64();
      float keep = 0.f;
      while (!stop) {
        uint32_t v[64];
        tc_ld32(tl, v);
        tc_ld32(tl + 32, v + 32);
        tc_wait_ld();
        if (mode & 8) {
#pragma unroll
          for (int j = 0; j < 64; ++j) keep += ex2_ftz(__uint_as_float(v[j]) * 1e-3f);
        } else {
          keep += __uint_as_float(v[lane & 63]);
        }
        next += period;
        while (clock64() < next && !stop) {}
      }
      if (keep == 123.456f) out[1] = 1;
    }
  } else if (warp >= 8 && warp < 12) {
    if (mode & 2) {
      const int r = (warp - 8) * 32 + lane;
      long long next = clock64();
      uint32_t x = threadIdx.x;
      int b = 0;
      while (!stop) {
        uint8_t* prow = sm + b * 16384 + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(0x3c003c00u, x, 0x3c003c00u, 0x3c003c00u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        b ^= 1;
        next += period;
        while (clock64() < next && !stop) {}
      }
    }
  } else if (warp == 12) {
    if ((mode & 4) && lane == 0) {
      long long next = clock64();
      uint32_t ph = 0;
      int st = 0;
      while (!stop) {
        mbar_expect_tx(&lbar, 32768);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sbase + 32768 + 0 * st),
                     "l"(gsrc + (size_t)blockIdx.x * 65536 + (size_t)st * 32768), "r"(32768), "r"(smem_u32(&lbar))
                     : "memory");
        mbar_wait(&lbar, ph);
        ph ^= 1;
        st ^= 1;
        next += period;
        while (clock64() < next && !stop) {}
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 196608 + 1024;
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Named { const char* name; Case c; };
  std::vector<Named> cases = {
      {"TS  N64  B mn-major (S GEMM shape)", {0, 64, 1, 16, 32, 0}},
      {"TS  N128 B mn-major", {0, 128, 1, 16, 32, 0}},
      {"TS  N256 B mn-major", {0, 256, 1, 16, 32, 0}},
      {"TS  N64  B k-major", {0, 64, 0, 16, 32, 0}},
      {"TS  N128 B k-major", {0, 128, 0, 16, 32, 0}},
      {"TS  N256 B k-major", {0, 256, 0, 16, 32, 0}},
      {"SS  N64  k-major", {1, 64, 0, 16, 32, 0}},
      {"SS  N128 k-major", {1, 128, 0, 16, 32, 0}},
      {"SS  N256 k-major (O GEMM shape)", {1, 256, 0, 16, 32, 0}},
      {"tile loop: 16 x TS N64 + 4 x SS N256 per tile", {0, 0, 0, 20, 32, 1}},
  };
  for (int grid : {1, 148}) {
    printf("grid %d\n", grid);
    for (auto& nc : cases) {
      long long h[2] = {0, 0};
      mma_bench_kernel<<<grid, 128, smem>>>(nc.c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", nc.name, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const int n_mma = nc.c.per_group * nc.c.groups;
      const double ideal = nc.c.interleave ? 1024.0 * nc.c.groups : (double)n_mma * nc.c.n / 2.0;
      printf("  %-48s %4d MMAs: issue %7lld cyc, total %7lld cyc = %6.1f cyc/MMA (ideal %5.1f) -> %.0f %% of the tensor roof\n", nc.name, n_mma,
             h[0], h[1], (double)h[1] / n_mma, ideal / n_mma, 100.0 * ideal / h[1]);
    }
  }
  uint8_t* gsrc;
  cudaMalloc(&gsrc, (size_t)148 * 65536);
  cudaMemset(gsrc, 0x3c, (size_t)148 * 65536);
  cudaFuncSetAttribute(mma_stress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 256;
  printf("tile loop (%d tiles, ideal %d cycles) beside the kernel's other traffic, grid 148\n", tiles, tiles * 1024);
  for (int grid : {148}) for (int period : {1024}) {
    for (int mode : {1024, 1024 + 512, 1024 + 128, 1024 + 256, 1024 + 128 + 15, 1024 + 256 + 15}) {
      long long h[2] = {0, 0};
      mma_stress_kernel<<<grid, 448, smem>>>(mode, period, tiles, gsrc, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("stress mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("  grid %3d burst period %4d  [%s%s%s%s]%*s %7lld cyc = %6.1f per tile -> %.0f %% of the tensor roof\n", grid, period, (mode & 1) || (mode & 8) ? " tcgen05.ld" : "",
             (mode & 8) ? "+ex2" : "", (mode & 2) ? " P-stores" : "", (mode & 4) ? " bulk-loads" : "", 2, (mode & 1024) ? ((mode & 128) ? " one burst/tile + 3 commits, 3-barrier probe + fence between bursts" : (mode & 256) ? " one burst/tile + 3 commits, flag poll + fence between bursts" : (mode & 512) ? " one burst/tile + 3 commits, fence between bursts" : " one burst/tile + 3 commits") : "", h[0], (double)h[0] / tiles,
             100.0 * tiles * 1024 / h[0]);
    }
  }
  return 0;
}
