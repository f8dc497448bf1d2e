// Probe for the integer tensor-core path on sm_100a: tcgen05.mma kind::i8 (S8/U8 x S8 -> S32 in TMEM) with
// K-major, non-swizzled ("interleaved" core-matrix) shared-memory operands, the layout k_fwd3_ozaki uses.
//   1. correctness of the shared-memory / instruction descriptors and of the TMEM accumulator layout
//      (D[128 x N] = A[128 x 64] * B[N x 64]^T against a host loop), signed and unsigned A, accumulate flag
//   2. issue-rate: back-to-back MMAs per SM -> integer MAC/clk/SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); LBO = distance between core
// matrices adjacent in K, SBO = distance between 8-row groups (cute/atom/mma_traits_sm100.hpp, Major-K INTERLEAVE)
__host__ __device__ inline uint32_t kmajor_offset(int r, int k, int lbo, int sbo) {
  return (r >> 3) * sbo + (k >> 4) * lbo + (r & 7) * 16 + (k & 15);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_signed, int b_signed) {
  uint32_t d = 0;
  d |= 2u << 4;                       // c_format = S32
  d |= (uint32_t)(a_signed & 1) << 7; // a_format: 0 = U8, 1 = S8
  d |= (uint32_t)(b_signed & 1) << 10;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;                           // K-major A and B, dense, no saturate
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait: returns false on timeout (a wrong descriptor must not hang the GPU)
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long max_clk) {
  long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > max_clk) return false;
  }
}

template <int N>
__global__ void __launch_bounds__(128) k_probe(const int8_t* A, const int8_t* B, int32_t* D, int a_signed, int* status) {
  __shared__ __align__(128) uint8_t sa[128 * 64];
  __shared__ __align__(128) uint8_t sb[N * 64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int LBO = 128, SBO = 512;
  for (int i = tid; i < 128 * 64; i += 128) sa[kmajor_offset(i / 64, i % 64, LBO, SBO)] = (uint8_t)A[i];
  for (int i = tid; i < N * 64; i += 128) sb[kmajor_offset(i / 64, i % 64, LBO, SBO)] = (uint8_t)B[i];
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy smem writes -> visible to the async proxy (tensor core)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, a_signed, 1);
    // K = 64 = two instructions of K = 32 (two 16-byte core-matrix columns each); second pass accumulates again
    for (int rep = 0; rep < 2; ++rep)
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * LBO, LBO, SBO);
        const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * LBO, LBO, SBO);
        umma_i8(tmem, ad, bd, idesc, (rep | ks) ? 1u : 0u);
      }
    umma_commit(&bar);
  }
  const bool ok = mbar_wait_bounded(&bar, 0, 200000000LL);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    // thread = row: warp w reads TMEM lanes 32w .. 32w+31
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// issue-rate: one CTA per SM, `iters` groups of `per` MMAs, each group committed and waited for
template <int N>
__global__ void __launch_bounds__(128) k_rate(int iters, int per, long long* clocks, int* status) {
  __shared__ __align__(128) uint8_t sa[128 * 64];
  __shared__ __align__(128) uint8_t sb[N * 64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int LBO = 128, SBO = 512;
  for (int i = tid; i < 128 * 64; i += 128) sa[i] = (uint8_t)(i * 7);
  for (int i = tid; i < N * 64; i += 128) sb[i] = (uint8_t)(i * 13);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, 1, 1);
    const uint64_t ad = make_desc(smem_u32(sa), LBO, SBO), bd = make_desc(smem_u32(sb), LBO, SBO);
    long long t0 = clock64();
    bool ok = true;
    for (int it = 0; it < iters && ok; ++it) {
      for (int j = 0; j < per; ++j) umma_i8(tmem + (uint32_t)((j % 6) * N), ad, bd, idesc, j >= 6 ? 1u : 0u);
      umma_commit(&bar);
      ok = mbar_wait_bounded(&bar, (uint32_t)(it & 1), 200000000LL);
    }
    long long t1 = clock64();
    if (!ok) *status = 2;
    clocks[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N>
static int check(int a_signed) {
  std::vector<int8_t> A(128 * 64), B(N * 64);
  srand(7 + N + a_signed);
  for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB; int32_t* dD; int* dS;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dD, 0xff, 128 * N * 4));
  k_probe<N><<<1, 128>>>(dA, dB, dD, a_signed, dS);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D(128 * N); int st;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < N; ++n) {
      long long s = 0;
      for (int k = 0; k < 64; ++k) {
        int a = a_signed ? (int)A[i * 64 + k] : (int)(uint8_t)A[i * 64 + k];
        s += (long long)a * (int)B[n * 64 + k];
      }
      s *= 2;   // the kernel accumulates the product twice
      if (D[i * N + n] != (int32_t)s) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %lld\n", i, n, D[i * N + n], s); ++bad; }
    }
  printf("probe N=%d a_signed=%d status=%d mismatches=%d\n", N, a_signed, st, bad);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return bad || st;
}

template <int N>
static void rate(int sms, double ghz) {
  long long* dC; int* dS;
  CK(cudaMalloc(&dC, sms * 8)); CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4));
  const int iters = 200, per = 48;
  k_rate<N><<<sms, 128>>>(iters, per, dC, dS);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_rate<N><<<sms, 128>>>(iters, per, dC, dS);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(sms); int st;
  CK(cudaMemcpy(c.data(), dC, sms * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double macs = (double)iters * per * 128.0 * N * 32.0;
  printf("rate N=%d status=%d: %.1f clk per MMA (M128 N%d K32), %.0f MAC/clk/SM, all-SM %.1f Tera-MAC/s (kernel %.3f ms)\n", N, st,
         (double)c[0] / (iters * per), N, macs / (double)c[0], macs * sms / (ms * 1e-3) / 1e12, ms);
  cudaFree(dC); cudaFree(dS);
}


// ---- A operand in TMEM (tcgen05.mma "TS" form): row i = lane i, 4 consecutive K bytes per 32-bit column
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

template <int N>
__global__ void __launch_bounds__(128) k_probe_ts(const int8_t* A, const int8_t* B, int32_t* D, int* status, int iters, int per,
                                                   long long* clocks) {
  __shared__ __align__(128) uint8_t sb[N * 64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int LBO = 128, SBO = 512;
  for (int i = tid; i < N * 64; i += 128) sb[kmajor_offset(i / 64, i % 64, LBO, SBO)] = (uint8_t)B[i];
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t tmem_a = tmem + 448;          // 16 columns: K = 64 bytes per row
  {
    // thread = row: write its 64 bytes as 16 packed words
    const int row = warp * 32 + lane;
    uint32_t w[16];
    for (int j = 0; j < 16; ++j) {
      uint32_t v = 0;
      for (int b = 0; b < 4; ++b) v |= (uint32_t)(uint8_t)A[row * 64 + 4 * j + b] << (8 * b);
      w[j] = v;
    }
    const uint32_t addr = tmem_a + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
                 "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  bool ok = true;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, 1, 1);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * LBO, LBO, SBO);
      umma_i8_ts(tmem, tmem_a + ks * 8, bd, idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  ok = mbar_wait_bounded(&bar, 0, 200000000LL);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // rate: back-to-back TS MMAs, 6 rotating accumulators
  if (tid == 0 && ok && iters > 0) {
    const uint32_t idesc = make_idesc(128, N, 1, 1);
    const uint64_t bd = make_desc(smem_u32(sb), LBO, SBO);
    long long t0 = clock64();
    for (int it = 0; it < iters && ok; ++it) {
      constexpr int NACC = (448 / N) < 6 ? (448 / N) : 6;
      for (int j = 0; j < per; ++j) umma_i8_ts(tmem + (uint32_t)((j % NACC) * N), tmem_a, bd, idesc, j >= NACC ? 1u : 0u);
      umma_commit(&bar);
      ok = mbar_wait_bounded(&bar, (uint32_t)((it + 1) & 1), 200000000LL);
    }
    long long t1 = clock64();
    if (!ok) *status = 2;
    clocks[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N>
static int check_ts(int sms) {
  std::vector<int8_t> A(128 * 64), B(N * 64);
  srand(11 + N);
  for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB; int32_t* dD; int* dS; long long* dC;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMalloc(&dC, sms * 8));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dD, 0xff, 128 * N * 4));
  const int iters = 200, per = 48;
  k_probe_ts<N><<<1, 128>>>(dA, dB, dD, dS, 0, per, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D(128 * N); int st;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < N; ++n) {
      long long s = 0;
      for (int k = 0; k < 64; ++k) s += (long long)A[i * 64 + k] * (int)B[n * 64 + k];
      if (D[i * N + n] != (int32_t)s) { if (bad < 5) printf("  TS mismatch (%d,%d): got %d want %lld\n", i, n, D[i * N + n], s); ++bad; }
    }
  printf("probe TS (A in TMEM) N=%d status=%d mismatches=%d\n", N, st, bad);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_probe_ts<N><<<sms, 128>>>(dA, dB, dD, dS, iters, per, dC);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_probe_ts<N><<<sms, 128>>>(dA, dB, dD, dS, iters, per, dC);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(sms);
  CK(cudaMemcpy(c.data(), dC, sms * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double macs = (double)iters * per * 128.0 * N * 32.0;
  printf("rate TS N=%d status=%d: %.1f clk per MMA, %.0f MAC/clk/SM (kernel %.3f ms)\n", N, st, (double)c[0] / (iters * per),
         macs / (double)c[0], ms);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dC);
  return bad || st;
}

// ---- thread <-> element map of tcgen05.ld.16x256b (accumulator-fragment shaped loads)
__global__ void __launch_bounds__(128) k_ldmap(uint32_t* out) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  {
    uint32_t w[16];
    for (int j = 0; j < 16; ++j) w[j] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)j;   // (row << 8) | col
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
                 "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp: lanes 32*warp + {0, 16}: two 16-lane loads, columns 0..7 and 8..15
  for (int half = 0; half < 2; ++half)
    for (int cb = 0; cb < 2; ++cb) {
      uint32_t v[4];
      const uint32_t addr = tmem + ((uint32_t)(warp * 32 + half * 16) << 16) + cb * 8;
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 4; ++j) out[((tid * 2 + half) * 2 + cb) * 4 + j] = v[j];
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}
static void ldmap() {
  uint32_t* d; CK(cudaMalloc(&d, 128 * 16 * 4));
  k_ldmap<<<1, 128>>>(d);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> h(128 * 16);
  CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
  printf("tcgen05.ld.16x256b.x1 map (warp 1; lane: half/colblock -> (row,col) x4):\n");
  for (int lane = 0; lane < 32; ++lane) {
    int tid = 32 + lane;
    printf(" lane %2d:", lane);
    for (int half = 0; half < 2; ++half)
      for (int cb = 0; cb < 2; ++cb) {
        printf("  h%dc%d", half, cb);
        for (int j = 0; j < 4; ++j) { uint32_t v = h[((tid * 2 + half) * 2 + cb) * 4 + j]; printf(" (%u,%u)", v >> 8, v & 0xff); }
      }
    printf("\n");
  }
  cudaFree(d);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int bad = 0;
  bad |= check<64>(1); bad |= check<64>(0); bad |= check<32>(1); bad |= check<16>(0);
  bad |= check_ts<64>(prop.multiProcessorCount); bad |= check_ts<32>(prop.multiProcessorCount); bad |= check_ts<16>(prop.multiProcessorCount);
  bad |= check_ts<256>(prop.multiProcessorCount);
  rate<64>(prop.multiProcessorCount, 1.965); rate<32>(prop.multiProcessorCount, 1.965); rate<16>(prop.multiProcessorCount, 1.965);
  ldmap();
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
