// tc_probe_tmem_a.cu -- probe: tcgen05 kind::tf32 MMA with the A operand in TENSOR MEMORY (written by tcgen05.st from
// the threads that own the rows) instead of shared memory.  D[128 x N] = A[128 x 32] * B[N x 32]^T, 3xTF32 split.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe_tmem_a tc_probe_tmem_a.cu && ./tc_probe_tmem_a
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int M = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int* status) {
  extern __shared__ __align__(128) float dyn[];
  float (*sB)[N * K] = reinterpret_cast<float (*)[N * K]>(dyn);       // [hi / lo][chunk c][row group][8 rows][4]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < N) {
    for (int c = 0; c < K / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(B + tid * K + 4 * c);
      float4 hi = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
      float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
      const int off = c * (N * 4) + (tid >> 3) * 32 + (tid & 7) * 4;
      *reinterpret_cast<float4*>(&sB[0][off]) = hi;
      *reinterpret_cast<float4*>(&sB[1][off]) = lo;
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  // A row of this thread -> TMEM columns [32, 64) (hi) and [64, 96) (lo); D lives in columns [0, 32)
  {
    uint32_t hi[32], lo[32];
    for (int k = 0; k < K; ++k) {
      const float v = A[tid * K + k];
      const float h = tf32_round(v);
      hi[k] = __float_as_uint(h);
      lo[k] = __float_as_uint(v - h);
    }
    tmem_st32(lane_base + 32, hi);
    tmem_st32(lane_base + 64, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    if (elected) {
      int first = 1;
      const int terms[3][2] = {{1, 0}, {0, 1}, {0, 0}};     // (a part, b part): lo*hi, hi*lo, hi*hi
      for (int term = 0; term < 3; ++term) {
        for (int j = 0; j < K / 8; ++j) {
          const uint32_t a_tmem = tm + 32 + terms[term][0] * 32 + j * 8;
          const uint64_t db = make_desc(smem_u32(&sB[terms[term][1]][0]) + j * 2 * (N * 16), N * 16, 128);
          const uint32_t acc = first ? 0u : 1u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
              ::"r"(tm), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
          first = 0;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    __syncwarp();
  }
  {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    if (!done) { if (tid == 0) *status = 1; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; ++n) D[tid * N + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128));
}

template <int N>
int run() {
  float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hD = (float*)malloc(M * N * 4);
  srand(1234);
  for (int i = 0; i < M * K; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < N * K; ++i) hB[i] = ((float)rand() / RAND_MAX * 2.f - 1.f) * 3.f;
  float *dA, *dB, *dD;
  int* dS;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, M * N * 4);
  cudaMemset(dS, 0, 4);
  probe_kernel<N><<<1, 128, 2 * N * K * 4>>>(dA, dB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0, maxerr32 = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * (double)hB[n * K + k];
      float r32 = 0.f;
      for (int k = 0; k < K; ++k) r32 = fmaf(hA[m * K + k], hB[n * K + k], r32);
      if (fabs(ref - (double)r32) > maxerr32) maxerr32 = fabs(ref - (double)r32);
      double err = fabs(ref - (double)hD[m * N + n]);
      if (!(err <= maxerr)) maxerr = err;
      if (fabs(ref) > maxref) maxref = fabs(ref);
    }
  printf("A in TMEM, N=%d: cuda=%s status=%d max abs err %.3e (fp32 fma chain: %.3e; max |ref| %.3f)\n", N,
         cudaGetErrorString(e), st, maxerr, maxerr32, maxref);
  return (e != cudaSuccess) || st || !(maxerr < 1e-4);
}

int main() {
  int bad = 0;
  bad |= run<32>();
  bad |= run<16>();
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
