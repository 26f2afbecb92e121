// dqn_kernels.cu -- DQN update on the device: TD target, loss, hand-written backward pass of the GAT
// Q-network, deterministic gradient reduction, gradient clipping + Adam (train_gcn_dqn.py:112-137).
//
// dqn_grad_kernel keeps the env-tile scheme of the rollout (thread = node, CTA = floor(128/N) sampled
// transitions).  Per transition it runs the target network on s', the online network on s (keeping the
// activations), the TD error and the backward pass.  Everything a node contributes to another node's
// gradient goes through shared memory as a *gather* over per-env dense matrices (sum of alpha / d-logit per
// (target, source) pair), and all weight gradients are tile GEMMs  dW = L^T R  over the CTA's 128 node
// rows; per-CTA partials are then summed in CTA order by dqn_reduce_kernel.  No atomics anywhere, so the
// gradient is bit-reproducible from run to run.
#include "tile_device.cuh"

namespace swarm {

constexpr int kPartialStride = 1680;   // 1673 gradients + [1673] = sum of squared TD errors, padded
constexpr int kXPad = 8;

struct DqnParams {
  SwarmConfig cfg;
  const float* w_online;
  const float* w_target;
  SwarmReplay batch;
  const int64_t* indices;
  int32_t n_graphs;
  float gamma;
  float loss_scale;
  float* partials;
  float* td;
  int32_t epb, maxdeg;
  const SwarmTrainCtl* ctl;   // optional device-resident training cursor: skip when !ctl->updating
};

struct DqnLayout {
  int w_on, w_tg, st, h, asrc, wt, wd, inl, kv, ki, nbr, u, r, dp, dob, dh, x, dq, ds, dt, ma, mz, red, total;
};

__host__ __device__ inline DqnLayout dqn_layout(int n, int k, int maxdeg, int epb, int graph_mode) {
  const int T = kTileThreads;
  DqnLayout L;
  int off = 0;
  const bool knn = graph_mode == SWARM_GRAPH_KNN;
  auto take = [&](int bytes) { int o = off; off = tile_align16(off + bytes); return o; };
  L.w_on = take(TW_COUNT * 4);
  L.w_tg = take(TW_COUNT * 4);
  L.st = take(T * 16);
  L.h = take(T * kHPad * 4);
  L.asrc = take(T * 4);
  L.wt = take(maxdeg * T * 4);
  L.wd = take(maxdeg * T * 4);
  L.inl = take(maxdeg * T);
  L.kv = take(knn ? n * T * 4 : 0);
  L.ki = take(knn ? n * T : 0);
  L.nbr = take(knn ? k * T : 0);
  L.u = take(T * kHPad * 4);
  L.r = take(T * kHPad * 4);
  L.dp = take(T * kHPad * 4);
  L.dob = take(T * kHPad * 4);
  L.dh = take(T * kHPad * 4);
  L.x = take(T * kXPad * 4);
  L.dq = take(T * kW2Pad * 4);
  L.ds = take(T * 4);
  L.dt = take(T * 4);
  L.ma = take(epb * n * n * 4);
  L.mz = take(epb * n * n * 4);
  L.red = take(T * 4);
  L.total = off;
  return L;
}

__device__ __forceinline__ void store_row32(float* tile, int row, const float (&v)[32]) {
  float4* r = reinterpret_cast<float4*>(tile + row * kHPad);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) r[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
}

// acc[0..3] += sum_n A[n*lda + ac] * B[n*ldb + bc .. bc+3]   (A == nullptr: column of ones)
__device__ __forceinline__ float4 tile_gemm4(const float* A, int lda, int ac, const float* B, int ldb, int bc) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int n = 0; n < kTileThreads; ++n) {
    const float a = A ? A[n * lda + ac] : 1.0f;
    const float4 b = *reinterpret_cast<const float4*>(B + n * ldb + bc);
    acc.x = fmaf(a, b.x, acc.x);
    acc.y = fmaf(a, b.y, acc.y);
    acc.z = fmaf(a, b.z, acc.z);
    acc.w = fmaf(a, b.w, acc.w);
  }
  return acc;
}

__global__ void __launch_bounds__(kTileThreads) dqn_grad_kernel(const __grid_constant__ DqnParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SwarmConfig& c = p.cfg;
  const int T = kTileThreads;
  const int N = c.n_agents;
  const int K = c.knn_k;
  const bool knn = c.graph_mode == SWARM_GRAPH_KNN;
  const TileThread t = tile_thread(N, p.epb, p.n_graphs);
  const int tid = t.tid;

  if (p.ctl && !p.ctl->updating) return;      // replay ring not filled yet (train:113-115)

  const DqnLayout L = dqn_layout(N, K, p.maxdeg, p.epb, c.graph_mode);
  float* sw_on = reinterpret_cast<float*>(smem + L.w_on);
  float* sw_tg = reinterpret_cast<float*>(smem + L.w_tg);
  float4* sst = reinterpret_cast<float4*>(smem + L.st);
  TileGraphSmem g;
  g.sh = reinterpret_cast<float*>(smem + L.h);
  g.sas = reinterpret_cast<float*>(smem + L.asrc);
  g.swt = reinterpret_cast<float*>(smem + L.wt);
  g.sin = smem + L.inl;
  g.skv = reinterpret_cast<float*>(smem + L.kv);
  g.ski = smem + L.ki;
  g.snbr = smem + L.nbr;
  float* swd = reinterpret_cast<float*>(smem + L.wd);
  float* tU = reinterpret_cast<float*>(smem + L.u);
  float* tR = reinterpret_cast<float*>(smem + L.r);
  float* tDP = reinterpret_cast<float*>(smem + L.dp);
  float* tDO = reinterpret_cast<float*>(smem + L.dob);
  float* tDH = reinterpret_cast<float*>(smem + L.dh);
  float* tX = reinterpret_cast<float*>(smem + L.x);
  float* tDQ = reinterpret_cast<float*>(smem + L.dq);
  float* sds = reinterpret_cast<float*>(smem + L.ds);
  float* sdt = reinterpret_cast<float*>(smem + L.dt);
  float* mA = reinterpret_cast<float*>(smem + L.ma);
  float* mZ = reinterpret_cast<float*>(smem + L.mz);
  float* sred = reinterpret_cast<float*>(smem + L.red);

  stage_weights(p.w_online, sw_on, tid, T);
  stage_weights(p.w_target, sw_tg, tid, T);

  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s;
  int act = 0;
  float rew = 0.0f;
  if (t.active) {
    const long long slot = p.indices ? p.indices[t.env] : t.env;
    const long long ri = slot * N + t.i;
    s = reinterpret_cast<const float4*>(p.batch.state)[ri];
    s2 = reinterpret_cast<const float4*>(p.batch.next_state)[ri];
    act = p.batch.actions[ri];
    rew = p.batch.rewards[ri];
  }

  int deg = 0;
  if (!knn && t.active) deg = tile_in_edges_complete(g, t, N);
  float agg[32];
  float adst;

  // ---------------- target network on s' : y = r + gamma * max_a Q_target(s')  (train:120-121) ------
  sst[tid] = s2;
  __syncthreads();
  if (knn) {
    tile_knn_rows(g, t, sst, s2, N, K);
    if (t.active) deg = tile_in_edges_knn(g, t, N, K);
  }
  float y = 0.0f;
  {
    const float x2[7] = {s2.x, s2.y, s2.z, s2.w, c.goal_x, c.goal_y, (float)t.i};
    tile_gat_conv(g, t, sw_tg, x2, deg, agg, adst);
    if (t.active) {
      float q2[9];
      gat_head(agg, sw_tg, q2);
      float qmax = q2[0];
#pragma unroll
      for (int a = 1; a < 9; ++a) qmax = fmaxf(qmax, q2[a]);
      y = __fadd_rn(rew, __fmul_rn(p.gamma, qmax));
    }
  }
  __syncthreads();

  // ---------------- online network on s, activations kept ----------------------------------------
  sst[tid] = s;
  __syncthreads();
  if (knn) {
    tile_knn_rows(g, t, sst, s, N, K);
    if (t.active) deg = tile_in_edges_knn(g, t, N, K);
  }
  const float x[7] = {s.x, s.y, s.z, s.w, c.goal_x, c.goal_y, (float)t.i};
  tile_gat_conv(g, t, sw_on, x, deg, agg, adst);        // alpha_e in g.swt, h rows in g.sh, alpha_src in g.sas

  float u[32], r[32], dvec[32];
  float delta = 0.0f;
#pragma unroll
  for (int k = 0; k < 32; ++k) { u[k] = 0.f; r[k] = 0.f; dvec[k] = 0.f; }
  float dq = 0.0f;
  if (t.active) {
#pragma unroll
    for (int k = 0; k < 32; ++k) u[k] = agg[k];
    float q[9];
    gat_head_keep(u, r, sw_on, q);
    float v = q[0];
#pragma unroll
    for (int a = 1; a < 9; ++a) v = (act == a) ? q[a] : v;        // values = Q(s).gather(1, a)  (train:119)
    delta = __fsub_rn(v, y);
    dq = 2.0f * delta * p.loss_scale;                               // d mean((v - y)^2) / dv
    if (p.td) p.td[t.gidx] = delta;
  } else {
    // rows of inactive threads must be exact zeros: they take part in the tile GEMMs below
    float4* hrow = reinterpret_cast<float4*>(g.sh + tid * kHPad);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) hrow[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  sred[tid] = delta * delta;

  // publish forward tiles
  store_row32(tU, tid, u);
  store_row32(tR, tid, r);
  {
    float4* xr = reinterpret_cast<float4*>(tX + tid * kXPad);
    xr[0] = t.active ? make_float4(x[0], x[1], x[2], x[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    xr[1] = t.active ? make_float4(x[4], x[5], x[6], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    float* dqr = tDQ + tid * kW2Pad;
#pragma unroll
    for (int a = 0; a < kW2Pad; ++a) dqr[a] = (t.active && a == act) ? dq : 0.0f;
  }

  // ---------------- backward: lin2 -> ReLU -> lin1 -> tanh ------------------------------------------
  // dr = W2[a,:]^T dq ; dp = dr * [r > 0]
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float dr = sw_on[TW_W2T + k * kW2Pad + act] * dq;
    dvec[k] = r[k] > 0.0f ? dr : 0.0f;
  }
  store_row32(tDP, tid, dvec);
  // du[k] = sum_c W1[c][k] dp[c] ; do = du * (1 - u^2)
  {
    float dO[32];
    const float4* w1 = reinterpret_cast<const float4*>(sw_on + TW_W1T);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float acc = 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 w = w1[k * 8 + c4];
        acc = fmaf(w.x, dvec[4 * c4 + 0], acc);
        acc = fmaf(w.y, dvec[4 * c4 + 1], acc);
        acc = fmaf(w.z, dvec[4 * c4 + 2], acc);
        acc = fmaf(w.w, dvec[4 * c4 + 3], acc);
      }
      dO[k] = acc * (1.0f - u[k] * u[k]);
    }
    store_row32(tDO, tid, dO);
#pragma unroll
    for (int k = 0; k < 32; ++k) dvec[k] = dO[k];       // dvec = d(out_i) from here on
  }

  // ---------------- backward: aggregation + edge softmax + LeakyReLU ----------------------------------
  // zero this node's rows of the per-env (target, source) matrices
  if (t.active) {
    float* ra = mA + (t.el * N + t.i) * N;
    float* rz = mZ + (t.el * N + t.i) * N;
    for (int j = 0; j < N; ++j) { ra[j] = 0.0f; rz[j] = 0.0f; }
  }
  float dt_i = 0.0f;
  if (t.active) {
    // d alpha_e = <d out_i, h_j> ; softmax backward: d z_e = alpha_e (d alpha_e - sum_e' alpha_e' d alpha_e')
    float dot_sum = 0.0f;
    for (int e = 0; e < deg; ++e) {
      const int j = g.sin[e * T + tid];
      const float4* hj = reinterpret_cast<const float4*>(g.sh + (t.envbase + j) * kHPad);
      float da = 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 hv = hj[c4];
        da = fmaf(dvec[4 * c4 + 0], hv.x, da);
        da = fmaf(dvec[4 * c4 + 1], hv.y, da);
        da = fmaf(dvec[4 * c4 + 2], hv.z, da);
        da = fmaf(dvec[4 * c4 + 3], hv.w, da);
      }
      swd[e * T + tid] = da;
      dot_sum = fmaf(g.swt[e * T + tid], da, dot_sum);
    }
    float* ra = mA + (t.el * N + t.i) * N;
    float* rz = mZ + (t.el * N + t.i) * N;
    for (int e = 0; e < deg; ++e) {
      const int j = g.sin[e * T + tid];
      const float alpha = g.swt[e * T + tid];
      const float dz = alpha * (swd[e * T + tid] - dot_sum);
      const float raw = __fadd_rn(g.sas[t.envbase + j], adst);
      const float dzz = raw > 0.0f ? dz : 0.2f * dz;
      ra[j] += alpha;         // parallel edges (duplicates, double self loops) accumulate
      rz[j] += dzz;
      dt_i += dzz;
    }
  }
  sdt[tid] = dt_i;
  __syncthreads();            // tDO, mA, mZ complete

  // gather in the source role:  d h_j = sum_i A[i][j] d out_i + ds_j att_src + dt_j att_dst
  {
    float dh[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) dh[k] = 0.0f;
    float ds_j = 0.0f;
    if (t.active) {
      for (int ii = 0; ii < N; ++ii) {
        const float a = mA[(t.el * N + ii) * N + t.i];
        ds_j += mZ[(t.el * N + ii) * N + t.i];
        const float4* dr = reinterpret_cast<const float4*>(tDO + (t.envbase + ii) * kHPad);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 dv = dr[c4];
          dh[4 * c4 + 0] = fmaf(a, dv.x, dh[4 * c4 + 0]);
          dh[4 * c4 + 1] = fmaf(a, dv.y, dh[4 * c4 + 1]);
          dh[4 * c4 + 2] = fmaf(a, dv.z, dh[4 * c4 + 2]);
          dh[4 * c4 + 3] = fmaf(a, dv.w, dh[4 * c4 + 3]);
        }
      }
      const float* as = sw_on + TW_ATT_S;
      const float* ad = sw_on + TW_ATT_D;
#pragma unroll
      for (int k = 0; k < 32; ++k) dh[k] = fmaf(ds_j, as[k], fmaf(dt_i, ad[k], dh[k]));
    }
    store_row32(tDH, tid, dh);
    sds[tid] = ds_j;
  }
  __syncthreads();            // every tile complete

  // ---------------- weight gradients: tile GEMMs over the 128 node rows -------------------------------
  float* out = p.partials + (long long)blockIdx.x * kPartialStride;
  constexpr int G_W0 = 64, G_AS = 8, G_AD = 8, G_B0 = 8, G_W1 = 256, G_B1 = 8, G_W2 = 72, G_B2 = 3;
  constexpr int G_TOTAL = G_W0 + G_AS + G_AD + G_B0 + G_W1 + G_B1 + G_W2 + G_B2;
  for (int grp = tid; grp < G_TOTAL; grp += T) {
    int gi = grp;
    if (gi < G_W0) {                                   // dW0[c][k] = sum_n DH[n][c] X[n][k]
      const int cc = gi >> 1, k4 = (gi & 1) * 4;
      const float4 a = tile_gemm4(tDH, kHPad, cc, tX, kXPad, k4);
      float* o = out + SWARM_W_CONV_LIN + cc * 7 + k4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z;
      if (k4 == 0) o[3] = a.w;
      continue;
    }
    gi -= G_W0;
    if (gi < G_AS) {                                   // d att_src[c] = sum_n ds[n] H[n][c]
      const float4 a = tile_gemm4(sds, 1, 0, g.sh, kHPad, gi * 4);
      float* o = out + SWARM_W_ATT_SRC + gi * 4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
      continue;
    }
    gi -= G_AS;
    if (gi < G_AD) {                                   // d att_dst[c] = sum_n dt[n] H[n][c]
      const float4 a = tile_gemm4(sdt, 1, 0, g.sh, kHPad, gi * 4);
      float* o = out + SWARM_W_ATT_DST + gi * 4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
      continue;
    }
    gi -= G_AD;
    if (gi < G_B0) {                                   // d conv1.bias[c] = sum_n DO[n][c]
      const float4 a = tile_gemm4(nullptr, 0, 0, tDO, kHPad, gi * 4);
      float* o = out + SWARM_W_CONV_BIAS + gi * 4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
      continue;
    }
    gi -= G_B0;
    if (gi < G_W1) {                                   // dW1[c][k] = sum_n DP[n][c] U[n][k]
      const int cc = gi >> 3, k4 = (gi & 7) * 4;
      const float4 a = tile_gemm4(tDP, kHPad, cc, tU, kHPad, k4);
      float* o = out + SWARM_W_LIN1 + cc * 32 + k4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
      continue;
    }
    gi -= G_W1;
    if (gi < G_B1) {                                   // d lin1.bias[c] = sum_n DP[n][c]
      const float4 a = tile_gemm4(nullptr, 0, 0, tDP, kHPad, gi * 4);
      float* o = out + SWARM_W_LIN1_BIAS + gi * 4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
      continue;
    }
    gi -= G_B1;
    if (gi < G_W2) {                                   // dW2[a][k] = sum_n DQ[n][a] R[n][k]
      const int aa = gi >> 3, k4 = (gi & 7) * 4;
      const float4 a = tile_gemm4(tDQ, kW2Pad, aa, tR, kHPad, k4);
      float* o = out + SWARM_W_LIN2 + aa * 32 + k4;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
      continue;
    }
    gi -= G_W2;
    {                                                  // d lin2.bias[a] = sum_n DQ[n][a]
      const float4 a = tile_gemm4(nullptr, 0, 0, tDQ, kW2Pad, gi * 4);
      float* o = out + SWARM_W_LIN2_BIAS + gi * 4;
      o[0] = a.x;
      if (gi < 2) { o[1] = a.y; o[2] = a.z; o[3] = a.w; }
    }
  }
  if (tid == 0) {
    float sse = 0.0f;
    for (int n = 0; n < T; ++n) sse += sred[n];
    out[SWARM_W_COUNT] = sse;
  }
}

// grad[o] = sum over CTAs (in CTA order) of partials[cta][o]; loss = loss_scale * sum of squared TD errors
__global__ void __launch_bounds__(256) dqn_reduce_kernel(const float* __restrict__ partials, int n_ctas, float loss_scale,
                                                         float* __restrict__ grad, float* __restrict__ loss,
                                                         const SwarmTrainCtl* __restrict__ ctl) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o > SWARM_W_COUNT) return;
  if (ctl && !ctl->updating) return;
  float acc = 0.0f;
  for (int b = 0; b < n_ctas; ++b) acc += partials[(long long)b * kPartialStride + o];
  if (o < SWARM_W_COUNT) grad[o] = acc;
  else loss[0] = acc * loss_scale;
}

// ---- clip_grad_norm_ + Adam (+ optional target copy), one CTA ---------------------------------------
struct AdamParams {
  float* w;
  const float* grad;
  float* m;
  float* v;
  float* target;
  float* grad_norm;
  float lerp_w;          // 1 - beta1
  float beta2;
  float one_minus_beta2;
  float neg_step_size;   // -(lr / (1 - beta1^t))
  float bc2_sqrt;        // sqrt(1 - beta2^t)
  float eps;
  float max_norm;
  // device-driven variant (swarm_train_tick_apply): step-dependent scalars are derived from *ctl on the device and
  // the cursor is advanced at the end
  SwarmTrainCtl* ctl;
  double lr, beta1, beta2d;
  long long ring_capacity;
  int32_t num_envs;
  int32_t update_target_every;
};

__global__ void __launch_bounds__(256) adam_clip_kernel(AdamParams p) {
  __shared__ double ssq[8][256];
  __shared__ float s_coef;
  __shared__ float s_neg_step, s_bc2_sqrt;
  __shared__ int s_update, s_sync;
  const int tid = threadIdx.x;
  if (p.ctl) {
    if (tid == 0) {
      const long long step = p.ctl->opt_step + 1;
      const double bc1 = 1.0 - pow(p.beta1, (double)step);
      const double bc2 = 1.0 - pow(p.beta2d, (double)step);
      s_neg_step = (float)(-(p.lr / bc1));
      s_bc2_sqrt = (float)sqrt(bc2);
      s_update = p.ctl->updating;
      s_sync = ((p.ctl->tick + 1) % p.update_target_every) == 0;
    }
    __syncthreads();
    p.neg_step_size = s_neg_step;
    p.bc2_sqrt = s_bc2_sqrt;
    if (!s_sync) p.target = nullptr;
    if (!s_update) {
      if (tid == 0) {
        p.ctl->tick += 1;
        p.ctl->ring_cursor = (p.ctl->ring_cursor + p.num_envs) % p.ring_capacity;
        const long long sz = p.ctl->ring_size + p.num_envs;
        p.ctl->ring_size = sz < p.ring_capacity ? sz : p.ring_capacity;
      }
      return;
    }
  }
  const int seg_off[9] = {SWARM_W_CONV_LIN, SWARM_W_ATT_SRC, SWARM_W_ATT_DST, SWARM_W_CONV_BIAS, SWARM_W_LIN1,
                          SWARM_W_LIN1_BIAS, SWARM_W_LIN2, SWARM_W_LIN2_BIAS, SWARM_W_COUNT};
  // torch.nn.utils.clip_grad_norm_: norms = [||g_t||_2 for each parameter tensor]; total = ||norms||_2
  for (int sgi = 0; sgi < 8; ++sgi) {
    double acc = 0.0;
    for (int o = seg_off[sgi] + tid; o < seg_off[sgi + 1]; o += 256) {
      const double gval = (double)p.grad[o];
      acc += gval * gval;
    }
    ssq[sgi][tid] = acc;
  }
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (tid < stride)
      for (int sgi = 0; sgi < 8; ++sgi) ssq[sgi][tid] += ssq[sgi][tid + stride];
    __syncthreads();
  }
  if (tid == 0) {
    double tot = 0.0;
    for (int sgi = 0; sgi < 8; ++sgi) {
      const float nrm = (float)sqrt(ssq[sgi][0]);
      tot += (double)nrm * (double)nrm;
    }
    const float total_norm = (float)sqrt(tot);
    float coef = 1.0f;
    if (p.max_norm > 0.0f) {
      coef = p.max_norm / (total_norm + 1e-6f);
      coef = coef > 1.0f ? 1.0f : coef;
    }
    s_coef = coef;
    if (p.grad_norm) p.grad_norm[0] = total_norm;
  }
  __syncthreads();
  const float coef = s_coef;
  for (int o = tid; o < SWARM_W_COUNT; o += 256) {
    const float gval = __fmul_rn(p.grad[o], coef);
    // exp_avg.lerp_(grad, 1 - beta1)
    float m = p.m[o];
    m = __fadd_rn(m, __fmul_rn(p.lerp_w, __fsub_rn(gval, m)));
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    float v = __fmul_rn(p.v[o], p.beta2);
    v = __fadd_rn(v, __fmul_rn(__fmul_rn(p.one_minus_beta2, gval), gval));
    // denom = sqrt(v) / sqrt(bias_correction2) + eps ; param.addcdiv_(exp_avg, denom, value = -step_size)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), p.bc2_sqrt), p.eps);
    const float w = __fadd_rn(p.w[o], __fdiv_rn(__fmul_rn(p.neg_step_size, m), denom));
    p.m[o] = m;
    p.v[o] = v;
    p.w[o] = w;
    if (p.target) p.target[o] = w;
  }
  if (p.ctl && tid == 0) {
    p.ctl->tick += 1;
    p.ctl->ring_cursor = (p.ctl->ring_cursor + p.num_envs) % p.ring_capacity;
    const long long sz = p.ctl->ring_size + p.num_envs;
    p.ctl->ring_size = sz < p.ring_capacity ? sz : p.ring_capacity;
    p.ctl->opt_step += 1;
  }
}

// ---- replay sampling on the device (GraphReplayBuffer.sample's index draw, train:38-39) ------------------------
// indices[g] ~ U{0 .. size-1} with size = the ring fill AFTER this tick's push; counter RNG keyed by (seed, tick, g).
__global__ void __launch_bounds__(256) train_sample_kernel(SwarmTrainCtl* ctl, int num_envs, long long capacity, int G,
                                                           unsigned long long seed, int64_t* __restrict__ indices) {
  long long size = ctl->ring_size + num_envs;
  if (size > capacity) size = capacity;
  const unsigned long long tick = (unsigned long long)(ctl->tick + 1);
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G && size > 0) {
    const uint64_t r = rng_draw(seed, (uint64_t)g, tick, 0x5A17u);
    indices[g] = (int64_t)__umul64hi(r, (uint64_t)size);
  }
  if (g == 0) ctl->updating = (size >= G) ? 1 : 0;
}

// ---- launchers -------------------------------------------------------------------------------
int dqn_maxdeg(const SwarmConfig& c) {
  return c.graph_mode == SWARM_GRAPH_KNN ? (c.n_agents + c.knn_k + 1) : c.n_agents;
}

long long dqn_workspace_bytes(const SwarmConfig& c, int n_graphs) {
  const int epb = kTileThreads / c.n_agents;
  const long long ctas = (n_graphs + epb - 1) / epb;
  return ctas * kPartialStride * 4 + 256;
}

int dqn_smem_bytes(const SwarmConfig& c) {
  const int epb = kTileThreads / c.n_agents;
  return dqn_layout(c.n_agents, c.knn_k, dqn_maxdeg(c), epb, c.graph_mode).total;
}

cudaError_t launch_train_sample(SwarmTrainCtl* ctl, int num_envs, long long capacity, int G, unsigned long long seed,
                                int64_t* indices, cudaStream_t stream) {
  train_sample_kernel<<<(G + 255) / 256, 256, 0, stream>>>(ctl, num_envs, capacity, G, seed, indices);
  return cudaGetLastError();
}

cudaError_t launch_dqn_grad(const SwarmConfig& c, const float* w_online, const float* w_target, const SwarmReplay& batch,
                            const int64_t* indices, int n_graphs, float gamma, float loss_scale, float* grad, float* loss,
                            float* td, void* workspace, cudaStream_t stream, const SwarmTrainCtl* ctl) {
  DqnParams p;
  p.ctl = ctl;
  p.cfg = c;
  p.w_online = w_online;
  p.w_target = w_target;
  p.batch = batch;
  p.indices = indices;
  p.n_graphs = n_graphs;
  p.gamma = gamma;
  p.loss_scale = loss_scale;
  p.partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  p.td = td;
  p.epb = kTileThreads / c.n_agents;
  p.maxdeg = dqn_maxdeg(c);
  const int smem = dqn_smem_bytes(c);
  cudaError_t err = cudaFuncSetAttribute(dqn_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return err;
  const int ctas = (n_graphs + p.epb - 1) / p.epb;
  dqn_grad_kernel<<<ctas, kTileThreads, smem, stream>>>(p);
  dqn_reduce_kernel<<<(SWARM_W_COUNT + 1 + 255) / 256, 256, 0, stream>>>(p.partials, ctas, loss_scale, grad, loss, ctl);
  return cudaGetLastError();
}

cudaError_t launch_adam_clip(float* w, const float* grad, float* m, float* v, long long step, double lr, double beta1,
                             double beta2, double eps, double max_norm, float* target, float* grad_norm,
                             cudaStream_t stream, SwarmTrainCtl* ctl, int num_envs, long long ring_capacity,
                             int update_target_every) {
  AdamParams p;
  p.ctl = ctl;
  p.lr = lr;
  p.beta1 = beta1;
  p.beta2d = beta2;
  p.num_envs = num_envs;
  p.ring_capacity = ring_capacity;
  p.update_target_every = update_target_every > 0 ? update_target_every : 1;
  p.w = w;
  p.grad = grad;
  p.m = m;
  p.v = v;
  p.target = target;
  p.grad_norm = grad_norm;
  // scalars exactly as torch.optim.Adam computes them (Python doubles, cast to float where they meet tensors)
  const double b1 = beta1, b2 = beta2;
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  p.lerp_w = (float)(1.0 - b1);
  p.beta2 = (float)beta2;
  p.one_minus_beta2 = (float)(1.0 - b2);
  p.neg_step_size = (float)(-(lr / bc1));
  p.bc2_sqrt = (float)sqrt(bc2);
  p.eps = (float)eps;
  p.max_norm = (float)max_norm;
  adam_clip_kernel<<<1, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace swarm
