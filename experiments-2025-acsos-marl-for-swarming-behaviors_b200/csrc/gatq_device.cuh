// gatq_device.cuh -- per-node pieces of the GAT Q-network (src/training/train_gcn_dqn.py:50-70 on
// torch_geometric 2.5.3 GATConv), shared by the env-tile kernels and the generic CSR kernels.
// `sw` is the shared-memory weight block in the k-major layout of tile_kernels.cuh (TW_*).
//
// Dense contractions run as sequential-k FFMA chains from 0 with the bias added afterwards (the order a
// BLAS micro-kernel uses for these tiny shapes); element-wise steps that torch executes as separate ops
// (message = alpha * h_j, scatter-add, + bias) are kept as separately rounded operations.
#ifndef SWARM_GATQ_DEVICE_CUH
#define SWARM_GATQ_DEVICE_CUH

#include "tile_kernels.cuh"

namespace swarm {

// acc[0..3] += x * w  as two packed FFMA2 (sm_100): each lane is the same IEEE fused multiply-add as fmaf, so the
// results are bit-identical to four scalar FFMAs at half the issue slots
__device__ __forceinline__ void fma4_packed(float x, const float4& w, float& a0, float& a1, float& a2, float& a3) {
  const float2 xx = make_float2(x, x);
  const float2 lo = __ffma2_rn(xx, make_float2(w.x, w.y), make_float2(a0, a1));
  const float2 hi = __ffma2_rn(xx, make_float2(w.z, w.w), make_float2(a2, a3));
  a0 = lo.x; a1 = lo.y; a2 = hi.x; a3 = hi.y;
}

// h = x W0^T (GATConv.lin, no bias); alpha_src = <h, att_src>, alpha_dst = <h, att_dst>
__device__ __forceinline__ void gat_project(const float (&x)[7], const float* __restrict__ sw, float (&h)[32],
                                            float& asrc, float& adst) {
#pragma unroll
  for (int cc = 0; cc < 32; ++cc) h[cc] = 0.0f;
  const float4* w0 = reinterpret_cast<const float4*>(sw + TW_W0T);
#pragma unroll
  for (int k = 0; k < 7; ++k) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      fma4_packed(x[k], w0[k * 8 + c4], h[4 * c4 + 0], h[4 * c4 + 1], h[4 * c4 + 2], h[4 * c4 + 3]);
    }
  }
  asrc = 0.0f;
  adst = 0.0f;
  const float* as = sw + TW_ATT_S;
  const float* ad = sw + TW_ATT_D;
#pragma unroll
  for (int cc = 0; cc < 32; ++cc) {
    asrc = __fadd_rn(asrc, __fmul_rn(h[cc], as[cc]));
    adst = __fadd_rn(adst, __fmul_rn(h[cc], ad[cc]));
  }
}

// LeakyReLU(0.2) attention logit of edge j -> i
__device__ __forceinline__ float gat_logit(float asrc_j, float adst_i) {
  const float z = __fadd_rn(asrc_j, adst_i);
  return z > 0.0f ? z : __fmul_rn(z, 0.2f);
}

// out += alpha * h_j.  FUSED = false: message rounded, then accumulated (two roundings, exactly torch's
// `alpha * x_j` followed by scatter-add); FUSED = true: one FFMA per channel (half the instructions, one rounding).
template <bool FUSED = false>
__device__ __forceinline__ void gat_accumulate(float (&out)[32], float alpha, const float4* __restrict__ hj) {
  if (FUSED) {
    // packed FFMA2 (sm_100): two channels per issue slot
    const float2 a2 = make_float2(alpha, alpha);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 v = hj[c4];
      const float2 lo = __ffma2_rn(a2, make_float2(v.x, v.y), make_float2(out[4 * c4 + 0], out[4 * c4 + 1]));
      const float2 hi = __ffma2_rn(a2, make_float2(v.z, v.w), make_float2(out[4 * c4 + 2], out[4 * c4 + 3]));
      out[4 * c4 + 0] = lo.x; out[4 * c4 + 1] = lo.y;
      out[4 * c4 + 2] = hi.x; out[4 * c4 + 3] = hi.y;
    }
    return;
  }
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 v = hj[c4];
    out[4 * c4 + 0] = __fadd_rn(out[4 * c4 + 0], __fmul_rn(alpha, v.x));
    out[4 * c4 + 1] = __fadd_rn(out[4 * c4 + 1], __fmul_rn(alpha, v.y));
    out[4 * c4 + 2] = __fadd_rn(out[4 * c4 + 2], __fmul_rn(alpha, v.z));
    out[4 * c4 + 3] = __fadd_rn(out[4 * c4 + 3], __fmul_rn(alpha, v.w));
  }
}

// out += w * row, the row already in registers (packed FFMA2)
__device__ __forceinline__ void gat_accumulate_regs(float (&out)[32], float w, const float4 (&v)[8]) {
  const float2 a2 = make_float2(w, w);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float2 lo = __ffma2_rn(a2, make_float2(v[c4].x, v[c4].y), make_float2(out[4 * c4 + 0], out[4 * c4 + 1]));
    const float2 hi = __ffma2_rn(a2, make_float2(v[c4].z, v[c4].w), make_float2(out[4 * c4 + 2], out[4 * c4 + 3]));
    out[4 * c4 + 0] = lo.x; out[4 * c4 + 1] = lo.y;
    out[4 * c4 + 2] = hi.x; out[4 * c4 + 3] = hi.y;
  }
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU (ex2.approx + rcp.approx): absolute error ~1e-7, 6 instructions
// instead of ~15 for tanhf.  Used by the tensor-core path, whose activations feed a 3xTF32 contraction anyway.
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (aggregate + conv1.bias) -> tanh -> lin1 -> ReLU -> lin2; returns argmax (first maximum wins).
// On return a1 holds u = tanh(agg + b0) and a2 holds r = relu(W1 u + b1) (kept for the backward pass).
// FAST: tanh on the SFU (tanh_fast, absolute error ~1e-7) -- for the input-space forwards of large swarms, whose
// attention already runs on SFU exponentials; the bit-faithful paths keep tanhf.
template <bool FAST = false>
__device__ __forceinline__ int gat_head_keep(float (&a1)[32], float (&a2)[32], const float* __restrict__ sw, float (&q)[9]) {
  const float* b0 = sw + TW_B0;
#pragma unroll
  for (int cc = 0; cc < 32; ++cc) {
    const float v = __fadd_rn(a1[cc], b0[cc]);
    a1[cc] = FAST ? tanh_fast(v) : tanhf(v);
  }

#pragma unroll
  for (int cc = 0; cc < 32; ++cc) a2[cc] = 0.0f;
  const float4* w1 = reinterpret_cast<const float4*>(sw + TW_W1T);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      fma4_packed(a1[k], w1[k * 8 + c4], a2[4 * c4 + 0], a2[4 * c4 + 1], a2[4 * c4 + 2], a2[4 * c4 + 3]);
    }
  }
  const float* b1 = sw + TW_B1;
#pragma unroll
  for (int cc = 0; cc < 32; ++cc) a2[cc] = fmaxf(__fadd_rn(a2[cc], b1[cc]), 0.0f);

  float qq[kW2Pad];
#pragma unroll
  for (int a = 0; a < kW2Pad; ++a) qq[a] = 0.0f;
  const float4* w2 = reinterpret_cast<const float4*>(sw + TW_W2T);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
#pragma unroll
    for (int a4 = 0; a4 < 3; ++a4) {
      fma4_packed(a2[k], w2[k * 3 + a4], qq[4 * a4 + 0], qq[4 * a4 + 1], qq[4 * a4 + 2], qq[4 * a4 + 3]);
    }
  }
  const float* b2 = sw + TW_B2;
  float best = 0.0f;
  int action = 0;
#pragma unroll
  for (int a = 0; a < 9; ++a) {
    q[a] = __fadd_rn(qq[a], b2[a]);
    if (a == 0 || q[a] > best) {
      best = q[a];
      action = a;
    }
  }
  return action;
}

__device__ __forceinline__ int gat_head(float (&a1)[32], const float* __restrict__ sw, float (&q)[9]) {
  float a2[32];
  return gat_head_keep<false>(a1, a2, sw, q);
}
__device__ __forceinline__ int gat_head_fast(float (&a1)[32], const float* __restrict__ sw, float (&q)[9]) {
  float a2[32];
  return gat_head_keep<true>(a1, a2, sw, q);
}

}  // namespace swarm
#endif
