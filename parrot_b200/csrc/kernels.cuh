// Non-GEMM kernels of the Parrot hot path (sm_100a): attention window (K7),
// GRU backward pre-pass, operand packing, emitter costs (K10/K11), GMM sampling
// (K12), optimizer (K14), encoder recurrence (K2), and a strided SIMT GEMM for
// the handful of tiny products that are not worth a tensor-core launch.
#pragma once
#include "engine.cuh"

namespace pb {

#define SQRT_1_2PI_F 0.3989422917366028f

// =========================================================================
// Attention window forward, one decoder step (model.py:664-690 / 931-958).
// One CTA per batch row.  phi is summed over the A mixture components serially
// in component order and w over text positions serially in position order, with
// contraction disabled, so that results track the reference's elementwise
// multiply + axis-sum to the last ulp of expf (bit-exact argmax requirement).
// =========================================================================
struct AttnFwdArgs {
  int B, U, C, A, H, Np, Cp;
  int type;  // 0 graves, 1 softmax
  float eps, align, sharp, timing;
  const float* h1;      // [B][H]
  const float* wT;      // [3A][H]  h1_to_att weights transposed: alpha | beta | kappa
  const float* batt;    // [3A]
  const float* ctx;     // [B][U][C]
  const float* k_prev;  // [B][A]
  float* k_out;         // [B][A]
  float* w_out;         // [B][C]
  bf16* w_hi;           // [Np][Cp] or null
  bf16* w_lo;
  float* phi_out;       // [B][U]
  float* ab_out;        // [B][2A]  alpha | beta
  float* e_out;         // [B][3A]  exp(a_hat) (softmax: probabilities) | exp(b_hat) | exp(k_hat)
  float* hat;           // [B][3A] scratch: pre-activations when the projection runs as its own stage (or null)
  float* hat_part;      // [H/32][B][3A] scratch: K-sliced partial projections (persistent scan / stand-alone step)
};
constexpr int ATT_KS = 32;   // features per projection slice

// Executed by every thread of the CTA for batch row b; `sh` needs
// rup(H,4) + 2*rup(3A,4) + rup(U,4) + nwarps*C floats.  Ends with a CTA barrier (sh may be reused).
// Projection stage of the attention step as a grid-wide pass: one (batch row, output) dot product of length H
// per warp, all loads of a lane in flight at once.  hat[b][j] = h1[b] . wT[j] + batt[j].
__device__ __forceinline__ void attention_proj_body(const AttnFwdArgs& a) {
  const int lane = threadIdx.x & 31;
  const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nw = (int)((gridDim.x * blockDim.x) >> 5);
  const int n = a.B * 3 * a.A;
  for (int idx = gw; idx < n; idx += nw) {
    const int b = idx / (3 * a.A), j = idx - b * 3 * a.A;
    const float* __restrict__ hr = a.h1 + (long long)b * a.H;
    const float* __restrict__ wr = a.wT + (long long)j * a.H;
    float s = 0.0f;
    if ((a.H & 127) == 0) {
      for (int k = lane * 4; k < a.H; k += 128) {
        const float4 h4 = __ldcg(reinterpret_cast<const float4*>(hr + k));
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
        s += h4.x * w4.x + h4.y * w4.y + h4.z * w4.z + h4.w * w4.w;
      }
    } else {
      for (int k = lane; k < a.H; k += 32) s = fmaf(__ldcg(hr + k), __ldg(wr + k), s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) a.hat[idx] = s + __ldg(a.batt + j);
  }
}

// WIDE: the window sum keeps 16 x 128-bit loads per lane in flight (measured faster with 8-warp CTAs, stand-alone
// and inside the persistent scan; with the earlier 12-warp CTAs at the 168-register cap it was slower)
#ifndef PB_ATT_WIDE
#define PB_ATT_WIDE 1
#endif
template <bool WIDE>
__device__ __forceinline__ void attention_fwd_body(const AttnFwdArgs& a, const int b, float* sh,
                                                   const bool precomputed_hat = false) {
  const int A3p = (3 * a.A + 3) & ~3;  // sections padded to 16 bytes (float4 accesses below)
  float* sh_h = sh;                          // H
  float* sh_hat = sh_h + ((a.H + 3) & ~3);   // 3A
  float* sh_abk = sh_hat + A3p;              // 3A : alpha, beta, kappa
  float* sh_phi = sh_abk + A3p;              // U
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = a.A;
  // fetched before the projection so that its latency is off the dependent chain
  const float k_prev_reg = (tid < A) ? a.k_prev[(long long)b * A + tid] : 0.0f;
  if (precomputed_hat) {
    if (tid < 3 * A) sh_hat[tid] = __ldcg(a.hat + (long long)b * 3 * A + tid);
    __syncthreads();
  } else {
  for (int i = tid; i < a.H; i += blockDim.x) sh_h[i] = a.h1[(long long)b * a.H + i];
  __syncthreads();
  // h1 . Watt^T: 3A dot products of length H, four outputs per warp pass
  for (int j0 = warp * 4; j0 < 3 * A; j0 += (blockDim.x >> 5) * 4) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = lane; k < a.H; k += 32) {
      const float hv = sh_h[k];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (j0 + r < 3 * A) s4[r] = fmaf(hv, __ldg(a.wT + (long long)(j0 + r) * a.H + k), s4[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float sv = s4[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (lane == 0 && j0 + r < 3 * A) sh_hat[j0 + r] = sv + a.batt[j0 + r];
    }
  }
  __syncthreads();
  }
  if (tid < A) {
    float ea;
    if (a.type == 1) {
      float m = sh_hat[0];
      for (int i = 1; i < A; ++i) m = fmaxf(m, sh_hat[i]);
      float sum = 0.0f;
      for (int i = 0; i < A; ++i) sum += expf(sh_hat[i] - m);
      ea = expf(sh_hat[tid] - m) / sum;
    } else {
      ea = expf(sh_hat[tid]);
    }
    const float eb = expf(sh_hat[A + tid]);
    const float ek = expf(sh_hat[2 * A + tid]);
    const float alpha = ea + a.eps;
    const float beta = (a.sharp == 1.0f ? eb : eb * a.sharp) + a.eps;
    const float step = (a.timing == 1.0f) ? a.align * ek : (a.align * ek) / a.timing;
    const float kappa = k_prev_reg + step;
    sh_abk[tid] = alpha;
    sh_abk[A + tid] = beta;
    sh_abk[2 * A + tid] = kappa;
    a.k_out[(long long)b * A + tid] = kappa;
    a.ab_out[(long long)b * 2 * A + tid] = alpha;
    a.ab_out[(long long)b * 2 * A + A + tid] = beta;
    a.e_out[(long long)b * 3 * A + tid] = ea;
    a.e_out[(long long)b * 3 * A + A + tid] = eb;
    a.e_out[(long long)b * 3 * A + 2 * A + tid] = ek;
  }
  __syncthreads();
  for (int u = tid; u < a.U; u += blockDim.x) {
    const float uf = (float)u;
    float phi = 0.0f;
    for (int i = 0; i < A; ++i) {
      const float d = __fsub_rn(sh_abk[2 * A + i], uf);
      const float d2 = __fmul_rn(d, d);
      float term;
      if (a.type == 1) {
        const float t1 = __fmul_rn(sh_abk[i], sqrtf(sh_abk[A + i]));
        term = __fmul_rn(t1, expf(__fmul_rn(__fmul_rn(-0.5f, sh_abk[A + i]), d2)));
      } else {
        term = __fmul_rn(sh_abk[i], expf(__fmul_rn(-sh_abk[A + i], d2)));
      }
      phi = __fadd_rn(phi, term);
    }
    if (a.type == 1) phi = __fmul_rn(SQRT_1_2PI_F, phi);
    sh_phi[u] = phi;
    a.phi_out[(long long)b * a.U + u] = phi;
  }
  __syncthreads();
  // w[c] = sum_u phi[u] * ctx[b][u][c]: warp w reduces its slice of text positions with 128-bit loads
  // (all loads of a lane independent -> full memory-level parallelism), then the slices are summed in
  // warp order.  (The reference sums over u serially; the reordering moves w by ~1e-7 relative and
  // cannot affect argmax(phi), which is already fixed above.)
  float* sh_part = sh_phi + ((a.U + 3) & ~3);   // [nwarp][C]
  const float* cb = a.ctx + (long long)b * a.U * a.C;
  const int nwarp = blockDim.x >> 5;
  const int per = (a.U + nwarp - 1) / nwarp;
  const int u0 = warp * per, u1 = min(a.U, u0 + per);
  if (WIDE && (a.C & 3) == 0 && a.C <= 256) {
    // both 128-column halves of 8 text positions in flight per lane (16 x 128 bit loads before the first use)
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    for (int ub = u0; ub < u1; ub += 8) {
      float4 x[8][2];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = lane * 4 + h * 128;
          x[r][h] = (ub + r < u1 && c < a.C)
                        ? __ldg(reinterpret_cast<const float4*>(cb + (long long)(ub + r) * a.C + c))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (ub + r < u1) {
          const float p = sh_phi[ub + r];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            acc[h].x = fmaf(p, x[r][h].x, acc[h].x); acc[h].y = fmaf(p, x[r][h].y, acc[h].y);
            acc[h].z = fmaf(p, x[r][h].z, acc[h].z); acc[h].w = fmaf(p, x[r][h].w, acc[h].w);
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = lane * 4 + h * 128;
      if (c < a.C) *reinterpret_cast<float4*>(sh_part + (long long)warp * a.C + c) = acc[h];
    }
  } else if ((a.C & 3) == 0) {
    for (int c = lane * 4; c < a.C; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int u = u0; u < u1; ++u) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(cb + (long long)u * a.C + c));
        const float p = sh_phi[u];
        acc.x = fmaf(p, x.x, acc.x); acc.y = fmaf(p, x.y, acc.y);
        acc.z = fmaf(p, x.z, acc.z); acc.w = fmaf(p, x.w, acc.w);
      }
      *reinterpret_cast<float4*>(sh_part + (long long)warp * a.C + c) = acc;
    }
  } else {
    for (int c = lane; c < a.C; c += 32) {
      float acc = 0.0f;
      for (int u = u0; u < u1; ++u) acc = fmaf(sh_phi[u], cb[(long long)u * a.C + c], acc);
      sh_part[(long long)warp * a.C + c] = acc;
    }
  }
  __syncthreads();
  for (int c = tid; c < a.C; c += blockDim.x) {
    float w = 0.0f;
    for (int q = 0; q < nwarp; ++q) w += sh_part[(long long)q * a.C + c];
    a.w_out[(long long)b * a.C + c] = w;
    if (a.w_hi) {
      bf16 hh, ll;
      split_bf16(w, hh, ll);
      a.w_hi[(long long)b * a.Cp + c] = hh;
      a.w_lo[(long long)b * a.Cp + c] = ll;
    }
  }
  __syncthreads();
}


// ---- projection stage, K-sliced (persistent scan and stand-alone step): CTA `slice` owns ATT_KS consecutive
// features of h1 and produces hat_part[slice][b][j] = sum_{k in slice} h1[b][k] * wT[j][k] for all rows / outputs.
// The window stage adds the slices in slice order (deterministic).  sh: (B + 3A) * (ATT_KS + 4) floats.
__device__ __forceinline__ void attention_proj_slice(const AttnFwdArgs& a, const int slice, float* sh) {
  constexpr int LD = ATT_KS + 4;   // row stride in floats: 16-byte aligned rows, conflict-free 128-bit reads
  const int A3 = 3 * a.A, k0 = slice * ATT_KS;
  float* sh_h = sh;                 // [B][LD]
  float* sh_w = sh + a.B * LD;      // [3A][LD]
  const int tid = threadIdx.x, nt = blockDim.x;
  // global -> shared with cp.async: all requests of a thread are in flight together (a load -> store loop costs one
  // L2 round trip per iteration)
  for (int i = tid; i < a.B * (ATT_KS / 4); i += nt) {
    const int b = i / (ATT_KS / 4), ch = i % (ATT_KS / 4);
    cp_async16(sh_h + b * LD + 4 * ch, a.h1 + (long long)b * a.H + k0 + 4 * ch);
  }
  for (int i = tid; i < A3 * (ATT_KS / 4); i += nt) {
    const int j = i / (ATT_KS / 4), ch = i % (ATT_KS / 4);
    cp_async16(sh_w + j * LD + 4 * ch, a.wT + (long long)j * a.H + k0 + 4 * ch);
  }
  cp_async_wait_all();
  __syncthreads();
  float* out = a.hat_part + (long long)slice * a.B * A3;
  // thread <-> (row b, output group jg): outputs j = jg + 4 i
  for (int idx = tid; idx < a.B * 4; idx += nt) {
    const int b = idx >> 2, jg = idx & 3;
    const float4* hr = reinterpret_cast<const float4*>(sh_h + b * LD);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < ATT_KS / 4; ++k4) {
      const float4 hv = hr[k4];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = jg + 4 * i;
        if (j < A3) {
          const float4 wv = reinterpret_cast<const float4*>(sh_w + j * LD)[k4];
          acc[i] = fmaf(hv.x, wv.x, acc[i]); acc[i] = fmaf(hv.y, wv.y, acc[i]);
          acc[i] = fmaf(hv.z, wv.z, acc[i]); acc[i] = fmaf(hv.w, wv.w, acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = jg + 4 * i;
      if (j < A3) out[b * A3 + j] = acc[i];
    }
  }
  __syncthreads();
}

// ---- window stage: CTA (b, part) of `nparts` CTAs per batch row.  Every part recomputes the (cheap) mixture
// parameters and phi for the row -- identical bits -- and reduces its own slice of the context columns; part 0
// writes kappa / alpha / beta / phi.  hat comes from the K-sliced partials (nslices > 0) or from a.hat.
// sh: 2*rup(3A,4) + rup(U,4) + nwarps*Cs floats (Cs = C / nparts).
// Projection of ONE batch row by the whole CTA (single-launch attention step): thread t owns features 4t .. 4t+3 (+ 4
// blockDim per pass), forms its share of the 3A dot products with 128-bit loads of h1 and W_att^T that are all
// independent (batches of 10 in flight), a 31-shuffle transpose-reduce leaves output j on lane j of every warp, and
// the warps' partials are added in warp order.  hat_out[j] = h1[b] . wT[j] + batt[j], j < 3A <= 32.
__device__ __forceinline__ void attention_proj_row(const AttnFwdArgs& a, const int b, float* sh_red, float* hat_out) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int A3 = 3 * a.A;
  float pv[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) pv[j] = 0.0f;
  const float* hr = a.h1 + (long long)b * a.H;
  for (int k = tid * 4; k < a.H; k += blockDim.x * 4) {
    const float4 h4 = __ldcg(reinterpret_cast<const float4*>(hr + k));
#pragma unroll
    for (int j0 = 0; j0 < 30; j0 += 10) {
      float4 w4[10];
#pragma unroll
      for (int i = 0; i < 10; ++i)
        w4[i] = (j0 + i < A3) ? __ldg(reinterpret_cast<const float4*>(a.wT + (long long)(j0 + i) * a.H + k))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        float acc = pv[j0 + i];
        acc = fmaf(h4.x, w4[i].x, acc); acc = fmaf(h4.y, w4[i].y, acc);
        acc = fmaf(h4.z, w4[i].z, acc); acc = fmaf(h4.w, w4[i].w, acc);
        pv[j0 + i] = acc;
      }
    }
  }
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    const bool up = (lane & k) != 0;
#pragma unroll
    for (int i = 0; i < k; ++i) {
      const float send = up ? pv[i] : pv[i + k];
      const float keep = up ? pv[i + k] : pv[i];
      pv[i] = keep + __shfl_xor_sync(0xffffffffu, send, k);
    }
  }
  sh_red[warp * 32 + lane] = pv[0];     // lane j: this warp's partial of output j
  __syncthreads();
  if (tid < A3) {
    float s = 0.0f;
    for (int w = 0; w < nwarp; ++w) s += sh_red[w * 32 + tid];
    hat_out[tid] = s + __ldg(a.batt + tid);
  }
  // (the caller's next __syncthreads publishes hat_out)
}

// PROJ: the CTA computes the projection of its row itself (attention_proj_row) right after its first context rows
// have been requested; needs 3A <= 32, H % 4 == 0 and 8 * 32 more floats of shared memory behind the window's.
template <bool WIDE, bool PROJ = false>
__device__ __forceinline__ void attention_window_part(const AttnFwdArgs& a, const int b, const int part,
                                                      const int nparts, const int nslices, float* sh) {
  const int A3p = (3 * a.A + 3) & ~3;
  float* sh_hat = sh;               // 3A
  float* sh_abk = sh_hat + A3p;     // 3A : alpha, beta, kappa
  float* sh_phi = sh_abk + A3p;     // U
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = a.A;
  // the first block of context rows of this warp is requested before anything else: it does not depend on phi, and a
  // dependent global round trip costs 1-2 us inside the scan
  const int Cs = a.C / nparts, c_base = part * Cs;
  const float* cb = a.ctx + (long long)b * a.U * a.C + c_base;
  const int nwarp = blockDim.x >> 5;
  const int per = (a.U + nwarp - 1) / nwarp;
  const int u0 = warp * per, u1 = min(a.U, u0 + per);
  const bool wide = WIDE && (Cs & 3) == 0 && (a.C & 3) == 0;
  float4 x0[16];
  if (wide) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
      x0[r] = (u0 + r < u1 && lane * 4 < Cs)
                  ? __ldg(reinterpret_cast<const float4*>(cb + (long long)(u0 + r) * a.C + lane * 4))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float k_prev_reg = (tid < A) ? __ldcg(a.k_prev + (long long)b * A + tid) : 0.0f;
  if (PROJ) {
    float* sh_red = sh_phi + ((a.U + 3) & ~3) + nwarp * Cs;   // [nwarp][32] behind the window's own arrays
    attention_proj_row(a, b, sh_red, sh_hat);
  } else
  if (tid < 3 * A) {
    float s;
    if (nslices > 0) {
      const float* src = a.hat_part + (long long)b * 3 * A + tid;
      const long long sstride = (long long)a.B * 3 * A;
      s = 0.0f;
      for (int q0 = 0; q0 < nslices; q0 += 32) {   // 32 partials requested together, added in slice order
        float x[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] = (q0 + q < nslices) ? __ldcg(src + (q0 + q) * sstride) : 0.0f;
#pragma unroll
        for (int q = 0; q < 32; ++q) s += x[q];
      }
      s += __ldg(a.batt + tid);
    } else {
      s = __ldcg(a.hat + (long long)b * 3 * A + tid);
    }
    sh_hat[tid] = s;
  }
  __syncthreads();
  if (tid < A) {
    float ea;
    if (a.type == 1) {
      float m = sh_hat[0];
      for (int i = 1; i < A; ++i) m = fmaxf(m, sh_hat[i]);
      float sum = 0.0f;
      for (int i = 0; i < A; ++i) sum += expf(sh_hat[i] - m);
      ea = expf(sh_hat[tid] - m) / sum;
    } else {
      ea = expf(sh_hat[tid]);
    }
    const float eb = expf(sh_hat[A + tid]);
    const float ek = expf(sh_hat[2 * A + tid]);
    const float alpha = ea + a.eps;
    const float beta = (a.sharp == 1.0f ? eb : eb * a.sharp) + a.eps;
    const float step = (a.timing == 1.0f) ? a.align * ek : (a.align * ek) / a.timing;
    const float kappa = k_prev_reg + step;
    sh_abk[tid] = alpha;
    sh_abk[A + tid] = beta;
    sh_abk[2 * A + tid] = kappa;
    if (part == 0) {
      a.k_out[(long long)b * A + tid] = kappa;
      a.ab_out[(long long)b * 2 * A + tid] = alpha;
      a.ab_out[(long long)b * 2 * A + A + tid] = beta;
      a.e_out[(long long)b * 3 * A + tid] = ea;
      a.e_out[(long long)b * 3 * A + A + tid] = eb;
      a.e_out[(long long)b * 3 * A + 2 * A + tid] = ek;
    }
  }
  __syncthreads();
  for (int u = tid; u < a.U; u += blockDim.x) {
    const float uf = (float)u;
    float phi = 0.0f;
    for (int i = 0; i < A; ++i) {
      const float d = __fsub_rn(sh_abk[2 * A + i], uf);
      const float d2 = __fmul_rn(d, d);
      float term;
      if (a.type == 1) {
        const float t1 = __fmul_rn(sh_abk[i], sqrtf(sh_abk[A + i]));
        term = __fmul_rn(t1, expf(__fmul_rn(__fmul_rn(-0.5f, sh_abk[A + i]), d2)));
      } else {
        term = __fmul_rn(sh_abk[i], expf(__fmul_rn(-sh_abk[A + i], d2)));
      }
      phi = __fadd_rn(phi, term);
    }
    if (a.type == 1) phi = __fmul_rn(SQRT_1_2PI_F, phi);
    sh_phi[u] = phi;
    if (part == 0) a.phi_out[(long long)b * a.U + u] = phi;
  }
  __syncthreads();
  // w[c] for c in [c_base, c_base + Cs): warp w reduces its slice of text positions (same partition and order as
  // attention_fwd_body: the result is bit-identical to the one-CTA-per-row kernel)
  float* sh_part = sh_phi + ((a.U + 3) & ~3);   // [nwarp][Cs]
  if (wide) {
    for (int c = lane * 4; c < Cs; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ub = u0; ub < u1; ub += 16) {
        float4 x[16];
        const bool first = (c == lane * 4) && (ub == u0);
#pragma unroll
        for (int r = 0; r < 16; ++r)
          x[r] = first ? x0[r]
                       : ((ub + r < u1) ? __ldg(reinterpret_cast<const float4*>(cb + (long long)(ub + r) * a.C + c))
                                        : make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          if (ub + r < u1) {
            const float p = sh_phi[ub + r];
            acc.x = fmaf(p, x[r].x, acc.x); acc.y = fmaf(p, x[r].y, acc.y);
            acc.z = fmaf(p, x[r].z, acc.z); acc.w = fmaf(p, x[r].w, acc.w);
          }
        }
      }
      *reinterpret_cast<float4*>(sh_part + (long long)warp * Cs + c) = acc;
    }
  } else {
    for (int c = lane; c < Cs; c += 32) {
      float acc = 0.0f;
      for (int u = u0; u < u1; ++u) acc = fmaf(sh_phi[u], cb[(long long)u * a.C + c], acc);
      sh_part[(long long)warp * Cs + c] = acc;
    }
  }
  __syncthreads();
  for (int c = tid; c < Cs; c += blockDim.x) {
    float w = 0.0f;
    for (int q = 0; q < nwarp; ++q) w += sh_part[(long long)q * Cs + c];
    a.w_out[(long long)b * a.C + c_base + c] = w;
    if (a.w_hi) {
      bf16 hh, ll;
      split_bf16(w, hh, ll);
      a.w_hi[(long long)b * a.Cp + c_base + c] = hh;
      a.w_lo[(long long)b * a.Cp + c_base + c] = ll;
    }
  }
  __syncthreads();
}
// CTAs per batch row of the window stage for a grid of `ctas`: C must split into float4-aligned slices
__host__ __device__ __forceinline__ int attention_nparts(int B, int C, int ctas) {
  int n = 1;
  while (n < 4 && 2 * n * B <= ctas && C % (8 * n) == 0 && C / (2 * n) >= 64) n *= 2;
  return n;
}
__global__ void __launch_bounds__(256) attention_proj_slice_kernel(const AttnFwdArgs a) {
  extern __shared__ float sh[];
  attention_proj_slice(a, blockIdx.x, sh);
}
__global__ void __launch_bounds__(256) attention_window_kernel(const AttnFwdArgs a, const int nparts, const int nslices) {
  extern __shared__ float sh[];
  attention_window_part<true>(a, blockIdx.x / nparts, blockIdx.x % nparts, nparts, nslices, sh);
}
// the whole attention step in ONE launch: projection + window per (batch row, context-column part)
__global__ void __launch_bounds__(256) attention_step_kernel(const AttnFwdArgs a, const int nparts) {
  extern __shared__ float sh[];
  attention_window_part<true, true>(a, blockIdx.x / nparts, blockIdx.x % nparts, nparts, 0, sh);
}

__global__ void __launch_bounds__(256) attention_fwd_kernel(const AttnFwdArgs a, const int precomputed_hat) {
  extern __shared__ float sh[];
  attention_fwd_body<true>(a, blockIdx.x, sh, precomputed_hat != 0);
}
// stage 1 of the two-kernel attention step: all 3A*B projections, one dot product per warp
__global__ void __launch_bounds__(256) attention_proj_kernel(const AttnFwdArgs a) { attention_proj_body(a); }

// =========================================================================
// Attention window backward, one decoder step (training form: sharp = timing = 1).
// Consumes dw (total gradient wrt w_t), the carried d(kappa), writes the gradient
// wrt the three attention pre-activations (fp32 + planes for the weight grads),
// adds datt * Watt^T into dh1, and updates the carried d(kappa).
// =========================================================================
struct AttnBwdArgs {
  int B, U, C, A, H, Np, Ap;   // Ap = padded 3A (plane width)
  int type;
  float eps, align;
  const float* dw;      // [B][C]
  const float* ctx;     // [B][U][C]
  const float* ab;      // [B][2A]
  const float* e;       // [B][3A]
  const float* kappa;   // [B][A] kappa_t
  float* dk_carry;      // [B][A] in: d/d kappa_t from step t+1 ; out: d/d kappa_{t-1}
  const float* watt;    // [3A][H] transposed h1_to_att weights
  float* dh1;           // [B][H] accumulated
  float* datt;          // [B][3A] fp32
  bf16* datt_hi;        // [Np][Ap]
  bf16* datt_lo;
  unsigned long long* dbg;   // debug: [8] globaltimer milestones of batch row 0 (or null)
};
#define ADBG(i) do { if (a.dbg && b == 0 && threadIdx.x == 0) a.dbg[i] = gtime(); } while (0)

__device__ __forceinline__ void gru_bwd_pre_rows(const ScanCtx& c, const int layer, const int t, const int b0,
                                                 const int b1, const int worker, const int nworkers);
// pre != nullptr (persistent backward scan): the GRU backward pre-pass of layer 1, step pre_t, row b is fused into the
// last stage (dh1 of that row is complete exactly there), see gru_bwd_pre_rows for the arithmetic.
// Load scheduling matters more than arithmetic here (every dependent global round trip costs 1-2 us inside the
// scan): the first block of context rows is requested before anything else, the first 16 rows of the transposed
// projection before the reductions (they do not depend on their results), the rest together with the pre-pass operands.
__device__ __forceinline__ void attention_bwd_body(const AttnBwdArgs& a, const int b, float* sh,
                                                   const ScanCtx* pre = nullptr, const int pre_t = 0) {
  float* sh_dw = sh;                    // C
  float* sh_dphi = sh_dw + a.C;         // U
  float* sh_red = sh_dphi + a.U;        // 3A * 16 warps
  float* sh_datt = sh_red + 3 * a.A * 16;// 3A, then 7A of prefetched per-row vectors
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = a.A, nwarp = blockDim.x >> 5;
  const float* cb = a.ctx + (long long)b * a.U * a.C;
  const bool fast = (a.C & 3) == 0 && a.C <= 256;
  float4 x[8][2];
  ADBG(0);
  if (fast) {
    const int u = warp * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane * 4 + h * 128;
        x[r][h] = (u + r < a.U && c < a.C)
                      ? __ldg(reinterpret_cast<const float4*>(cb + (long long)(u + r) * a.C + c))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  }
  for (int i = tid; i < a.C; i += blockDim.x) sh_dw[i] = __ldcg(a.dw + (long long)b * a.C + i);
  // small per-row vectors needed after the reductions: fetch them now, off the dependent chain
  float* sh_small = sh_datt + 3 * A;   // [ab 2A | kappa A | e 3A | dk_carry A]
  for (int i = tid; i < 7 * A; i += blockDim.x) {
    float v;
    if (i < 2 * A) v = a.ab[(long long)b * 2 * A + i];
    else if (i < 3 * A) v = a.kappa[(long long)b * A + (i - 2 * A)];
    else if (i < 6 * A) v = a.e[(long long)b * 3 * A + (i - 3 * A)];
    else v = __ldcg(a.dk_carry + (long long)b * A + (i - 6 * A));
    sh_small[i] = v;
  }
  __syncthreads();
  ADBG(1);
  if (fast) {
    // 8 text positions per warp pass, all loads of a pass (<= 16 x 128 bit per lane) issued before the first use
    for (int u = warp * 8; u < a.U; u += nwarp * 8) {
      if (u != warp * 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = lane * 4 + h * 128;
            x[r][h] = (u + r < a.U && c < a.C)
                          ? __ldg(reinterpret_cast<const float4*>(cb + (long long)(u + r) * a.C + c))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
      }
      float4 d[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane * 4 + h * 128;
        d[h] = (c < a.C) ? *reinterpret_cast<const float4*>(sh_dw + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float sv = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          sv += d[h].x * x[r][h].x + d[h].y * x[r][h].y + d[h].z * x[r][h].z + d[h].w * x[r][h].w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        if (lane == 0 && u + r < a.U) sh_dphi[u + r] = sv;
      }
    }
  } else
  for (int u = warp * 4; u < a.U; u += nwarp * 4) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    if ((a.C & 3) == 0) {
      for (int c = lane * 4; c < a.C; c += 128) {
        const float4 d = *reinterpret_cast<const float4*>(sh_dw + c);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (u + r < a.U) {
            const float4 xx = __ldg(reinterpret_cast<const float4*>(cb + (long long)(u + r) * a.C + c));
            s4[r] += d.x * xx.x + d.y * xx.y + d.z * xx.z + d.w * xx.w;
          }
      }
    } else {
      for (int c = lane; c < a.C; c += 32)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (u + r < a.U) s4[r] = fmaf(sh_dw[c], cb[(long long)(u + r) * a.C + c], s4[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float sv = s4[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (lane == 0 && u + r < a.U) sh_dphi[u + r] = sv;
    }
  }
  // first half of the dh1 update's operands (this thread's first feature quad): requested now, consumed after the
  // reductions
  const bool vec = (a.H & 3) == 0;
  const int f0 = tid * 4;
  float4 prev0 = make_float4(0.f, 0.f, 0.f, 0.f), w4[16];
  if (vec && f0 < a.H) {
    prev0 = ldcg4(a.dh1 + (long long)b * a.H + f0);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      w4[j] = (j < 3 * A) ? ldg4(a.watt + (long long)j * a.H + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  ADBG(2);
  // per-component reductions over u: warp w owns components w, w + nwarp, ...; lane l sums u = l, l + 32, ...; one
  // shuffle tree per component (the earlier form ran the A components one after the other on U of the threads)
  for (int i = warp; i < A; i += nwarp) {
    const float al = sh_small[i];
    const float be = sh_small[A + i];
    const float ka = sh_small[2 * A + i];
    float da = 0.0f, db = 0.0f, dk = 0.0f;
    for (int u = lane; u < a.U; u += 32) {
      const float d = ka - (float)u;
      const float d2 = d * d;
      float g = sh_dphi[u];
      if (a.type == 1) {
        g *= SQRT_1_2PI_F;
        const float sb = sqrtf(be);
        const float ee = expf(-0.5f * be * d2);
        da += g * sb * ee;
        db += g * al * ee * (0.5f / sb - sb * 0.5f * d2);
        dk += g * al * sb * ee * (-be * d);
      } else {
        const float ee = expf(-be * d2);
        da += g * ee;
        db += g * al * ee * (-d2);
        dk += g * al * ee * (-2.0f * be * d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      da += __shfl_xor_sync(0xffffffffu, da, o);
      db += __shfl_xor_sync(0xffffffffu, db, o);
      dk += __shfl_xor_sync(0xffffffffu, dk, o);
    }
    if (lane == 0) {
      sh_red[(0 * A + i) * 16] = da;
      sh_red[(1 * A + i) * 16] = db;
      sh_red[(2 * A + i) * 16] = dk;
    }
  }
  __syncthreads();
  if (tid < A) {
    const float da = sh_red[(0 * A + tid) * 16];
    const float db = sh_red[(1 * A + tid) * 16];
    const float dk = sh_red[(2 * A + tid) * 16] + sh_small[6 * A + tid];
    const float ea = sh_small[3 * A + tid];
    const float eb = sh_small[4 * A + tid];
    const float ek = sh_small[5 * A + tid];
    float da_hat;
    if (a.type == 1) {
      float dot = 0.0f;
      for (int i = 0; i < A; ++i) dot += sh_red[(0 * A + i) * 16] * sh_small[3 * A + i];
      da_hat = ea * (da - dot);
    } else {
      da_hat = da * ea;
    }
    sh_datt[tid] = da_hat;
    sh_datt[A + tid] = db * eb;
    sh_datt[2 * A + tid] = dk * a.align * ek;
    a.dk_carry[(long long)b * A + tid] = dk;
  }
  __syncthreads();
  if (tid < 3 * A) {
    const float v = sh_datt[tid];
    a.datt[(long long)b * 3 * A + tid] = v;
    bf16 hh, ll;
    split_bf16(v, hh, ll);
    a.datt_hi[(long long)b * a.Ap + tid] = hh;
    a.datt_lo[(long long)b * a.Ap + tid] = ll;
  }
  ADBG(3);
  if (vec) {
    // dh1[b][f..f+3] += sum_j datt[j] * watt[j][f..f+3]; with `pre`, the row's GRU backward pre-pass follows at once
    const LayerBuf* L = pre ? &pre->L[0] : nullptr;
    for (int f = f0; f < a.H; f += blockDim.x * 4) {
      float4 prev = prev0;
      if (f != f0) {
        prev = ldcg4(a.dh1 + (long long)b * a.H + f);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          w4[j] = (j < 3 * A) ? ldg4(a.watt + (long long)j * a.H + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // second request block: the remaining projection rows and the pre-pass operands
      float4 w5[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        w5[j] = (16 + j < 3 * A) ? ldg4(a.watt + (long long)(16 + j) * a.H + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 pz = prev, pc = prev, phv = prev, pd0 = prev;
      long long pi = 0;
      if (L) {
        pi = ((long long)pre_t * pre->B + b) * a.H + f;
        pz = ldcg4(L->z + pi); pc = ldcg4(L->c + pi); phv = ldcg4(L->h + pi); pd0 = ldcg4(L->dh + pi);
      }
      float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < 3 * A) {
          const float dv = sh_datt[j];
          sacc.x = fmaf(dv, w4[j].x, sacc.x); sacc.y = fmaf(dv, w4[j].y, sacc.y);
          sacc.z = fmaf(dv, w4[j].z, sacc.z); sacc.w = fmaf(dv, w4[j].w, sacc.w);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (16 + j < 3 * A) {
          const float dv = sh_datt[16 + j];
          sacc.x = fmaf(dv, w5[j].x, sacc.x); sacc.y = fmaf(dv, w5[j].y, sacc.y);
          sacc.z = fmaf(dv, w5[j].z, sacc.z); sacc.w = fmaf(dv, w5[j].w, sacc.w);
        }
      }
      const float4 dh1 = make_float4(prev.x + sacc.x, prev.y + sacc.y, prev.z + sacc.z, prev.w + sacc.w);
      *reinterpret_cast<float4*>(a.dh1 + (long long)b * a.H + f) = dh1;
      if (L) {
        const int H = a.H, Hp = pre->Hp;
        float4 dh0, dac, dagz;
#define PB_PRE1(m)                                           \
        {                                                    \
          const float dh = dh1.m, z = pz.m, cc = pc.m;       \
          dh0.m = pd0.m + dh * (1.0f - z);                   \
          dac.m = (dh * z) * (1.0f - cc * cc);               \
          dagz.m = (dh * (cc - phv.m)) * z * (1.0f - z);     \
        }
        PB_PRE1(x) PB_PRE1(y) PB_PRE1(z) PB_PRE1(w)
#undef PB_PRE1
        *reinterpret_cast<float4*>(L->dh + pi) = dh0;
        float* dap = L->da + ((long long)pre_t * pre->B + b) * 3 * H;
        *reinterpret_cast<float4*>(dap + f) = dac;
        *reinterpret_cast<float4*>(dap + H + f) = dagz;
        const long long po = ((long long)pre_t * pre->Np + b) * (3 * Hp);
        uint2 hh, ll;
        split4(dac, hh, ll);
        *reinterpret_cast<uint2*>(L->da_hi + po + f) = hh;
        *reinterpret_cast<uint2*>(L->da_lo + po + f) = ll;
        split4(dagz, hh, ll);
        *reinterpret_cast<uint2*>(L->da_hi + po + Hp + f) = hh;
        *reinterpret_cast<uint2*>(L->da_lo + po + Hp + f) = ll;
      }
    }
  } else {
    for (int f = tid; f < a.H; f += blockDim.x) {
      float s = 0.0f;
      const float prev = a.dh1[(long long)b * a.H + f];
      for (int j = 0; j < 3 * A; ++j) s = fmaf(sh_datt[j], __ldg(a.watt + (long long)j * a.H + f), s);
      a.dh1[(long long)b * a.H + f] = prev + s;
    }
    if (pre) {
      __syncthreads();
      gru_bwd_pre_rows(*pre, 0, pre_t, b, b + 1, tid, blockDim.x);
    }
  }
  ADBG(4);
  __syncthreads();
  ADBG(5);
}

__global__ void __launch_bounds__(256) attention_bwd_kernel(const AttnBwdArgs a) {
  extern __shared__ float sh[];
  attention_bwd_body(a, blockIdx.x, sh);
}

// =========================================================================
// GRU backward pre-pass for one layer / step: everything that is elementwise in
// dh_t (see oracle _gru_bwd).  Produces da_c and the update half of da_g.
// =========================================================================
struct PreArgs { int layer[3]; int t[3]; int n; };
__device__ __forceinline__ void gru_bwd_pre_body(const ScanCtx& c, const int layer, const int t) {
  // snapshot the context (it lives in global memory; see EpiLocal in engine.cuh)
  const LayerBuf L = c.L[layer];
  const int H = c.H, B = c.B, Np = c.Np, Hp = c.Hp;
  const long long n = (long long)B * H;
  const float* __restrict__ zp = L.z + (long long)t * n;
  const float* __restrict__ cp = L.c + (long long)t * n;
  const float* __restrict__ hp = L.h + (long long)t * n;
  float* __restrict__ dhp = L.dh + (long long)t * n;         // slot t ; slot t + 1 = + n
  float* __restrict__ dap = L.da + (long long)t * B * 3 * H;
  bf16* __restrict__ phi = L.da_hi + (long long)t * Np * (3 * Hp);
  bf16* __restrict__ plo = L.da_lo + (long long)t * Np * (3 * Hp);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), f = (int)(i % H);
    const float dh = dhp[i + n];  // slot t + 1
    const float z = zp[i], cc = cp[i], hpv = hp[i];
    const float dc = dh * z;
    const float dz = dh * (cc - hpv);
    dhp[i] += dh * (1.0f - z);
    const float dac = dc * (1.0f - cc * cc);
    const float dagz = dz * z * (1.0f - z);
    const long long ao = (long long)b * 3 * H;
    dap[ao + f] = dac;
    dap[ao + H + f] = dagz;
    const long long po = (long long)b * (3 * Hp);
    bf16 hh, ll;
    split_bf16(dac, hh, ll);
    phi[po + f] = hh;
    plo[po + f] = ll;
    split_bf16(dagz, hh, ll);
    phi[po + Hp + f] = hh;
    plo[po + Hp + f] = ll;
  }
}
// same as gru_bwd_pre_body restricted to batch rows [b0, b1), strided over an arbitrary worker set.  One item = four
// consecutive features of one row (128-bit accesses); two items of a thread are loaded before either is used.
struct PreItem { float4 dh1, z, c, h, dh0; };
__device__ __forceinline__ void gru_bwd_pre_rows(const ScanCtx& c, const int layer, const int t, const int b0,
                                                 const int b1, const int worker, const int nworkers) {
  const LayerBuf L = c.L[layer];
  const int H = c.H, B = c.B, Np = c.Np, Hp = c.Hp, H4 = H >> 2;
  const long long n = (long long)B * H;
  const float* __restrict__ zp = L.z + (long long)t * n;
  const float* __restrict__ cp = L.c + (long long)t * n;
  const float* __restrict__ hp = L.h + (long long)t * n;
  float* __restrict__ dhp = L.dh + (long long)t * n;
  float* __restrict__ dap = L.da + (long long)t * B * 3 * H;
  bf16* __restrict__ phi = L.da_hi + (long long)t * Np * (3 * Hp);
  bf16* __restrict__ plo = L.da_lo + (long long)t * Np * (3 * Hp);
  const long long n4 = (long long)(b1 - b0) * H4;
  auto load = [&](long long i4, PreItem& it) {
    const long long i = (long long)b0 * H + i4 * 4;
    it.dh1 = ldcg4(dhp + i + n); it.z = ldcg4(zp + i); it.c = ldcg4(cp + i); it.h = ldcg4(hp + i); it.dh0 = ldcg4(dhp + i);
  };
  auto finish = [&](long long i4, const PreItem& it) {
    const long long i = (long long)b0 * H + i4 * 4;
    const int b = (int)(i / H), f = (int)(i % H);
    float4 dh0, dac, dagz;
#define PB_PRE1(m)                                             \
    {                                                          \
      const float dh = it.dh1.m, z = it.z.m, cc = it.c.m;      \
      dh0.m = it.dh0.m + dh * (1.0f - z);                      \
      dac.m = (dh * z) * (1.0f - cc * cc);                     \
      dagz.m = (dh * (cc - it.h.m)) * z * (1.0f - z);          \
    }
    PB_PRE1(x) PB_PRE1(y) PB_PRE1(z) PB_PRE1(w)
#undef PB_PRE1
    *reinterpret_cast<float4*>(dhp + i) = dh0;
    const long long ao = (long long)b * 3 * H;
    *reinterpret_cast<float4*>(dap + ao + f) = dac;
    *reinterpret_cast<float4*>(dap + ao + H + f) = dagz;
    const long long po = (long long)b * (3 * Hp);
    uint2 hh, ll;
    split4(dac, hh, ll);
    *reinterpret_cast<uint2*>(phi + po + f) = hh;
    *reinterpret_cast<uint2*>(plo + po + f) = ll;
    split4(dagz, hh, ll);
    *reinterpret_cast<uint2*>(phi + po + Hp + f) = hh;
    *reinterpret_cast<uint2*>(plo + po + Hp + f) = ll;
  };
  for (long long i0 = worker; i0 < n4; i0 += 2LL * nworkers) {
    const long long i1 = i0 + nworkers;
    PreItem a0, a1;
    load(i0, a0);
    if (i1 < n4) load(i1, a1);
    finish(i0, a0);
    if (i1 < n4) finish(i1, a1);
  }
}
__global__ void gru_bwd_pre_kernel(const ScanCtx* cp, const PreArgs pa) {
  // grid (blocks, n): blockIdx.y selects the (layer, t) pair; the body strides over gridDim.x blocks
  if ((int)blockIdx.y >= pa.n) return;
  gru_bwd_pre_body(*cp, pa.layer[blockIdx.y], pa.t[blockIdx.y]);
}

// =========================================================================
// Persistent scan kernels: ONE cooperative launch runs all T decoder steps (forward) / the whole reverse
// sweep (backward).  Per tick the phases of the layer wavefront are separated by grid barriers instead of
// kernel boundaries; the GEMM pipeline state (smem ring, TMEM accumulators) lives across phases.
// =========================================================================
constexpr int ATT_SMEM_BYTES = 24 * 1024;
#define STAMP(S, bar, k)                                                                                \
  do {                                                                                                  \
    if ((S).stamps && (int)(bar) < (S).stamp_bars)                                                      \
      (S).stamps[((size_t)blockIdx.x * (S).stamp_bars + (bar)) * 2 + (k)] = gtime();                    \
  } while (0)

// one GEMM phase of a persistent kernel: wait for the previous grid barrier where data produced by other
// CTAs is consumed (producer: TMA of activation planes; epilogue: stashes), run the roles, arrive.
// One GEMM phase of a persistent kernel.  DIR 1 / 2: forward / backward sweep.  `chunk`: plain chunk table (jobs read
// from global memory) instead of a scan table (this CTA's job and epilogue context cached in shared memory).  There
// is ONE call site per kernel (the phases share the code; EngineParams live in shared memory): the persistent loop
// must stay small enough for the instruction caches.
// `ctr` / `ncta`: the barrier domain of the phase -- the whole grid in the wavefront kernels, one layer group in the
// grouped kernels.  `xctr` / `xtarget` (optional): a second, foreign counter that must reach `xtarget` before data of
// another group is touched (chunk products of the grouped kernels read the lower / upper layer's planes).
template <int DIR, class SP>
__device__ __forceinline__ void group_gemm_phase(Pipe& p, const EngineParams& P, int tick, const SP& S,
                                                 unsigned int* ctr, const int ncta, unsigned int& bar,
                                                 const PhaseCache* pc, const bool chunk,
                                                 const unsigned int* xctr = nullptr, const unsigned int xtarget = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int target = bar * (unsigned int)ncta;
  // the accumulator width of this phase (the smem ring keeps the stage stride chosen at kernel start)
  p.n_cols = P.n_cols;
  p.b_bytes = (uint32_t)P.n_cols * KB * 2;
  if (warp == 0) {
    if (lane == 0) {
      if (!chunk) {
        // scan phase: this CTA's cached job, weight tiles resident in tensor memory or prefetched ahead of the barrier
        producer_scan(p, P, tick, pc, ctr, target, S.prefetch != 0);
      } else {
        if (bar) grid_wait(ctr, target);
        if (xctr) grid_wait(xctr, xtarget);
        producer_run(p, P, tick, nullptr, 0, pc);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (!chunk) mma_scan(p, P, tick, pc);
    else mma_run(p, P, tick, pc);
  } else {
    if (warp == 2 && lane == 0) {   // one poller for the epilogue warps
      if (bar) grid_wait(ctr, target);
      if (xctr) grid_wait(xctr, xtarget);
    }
    epi_group_sync();
    if (threadIdx.x == 64) { STAMP(S, bar, 0); TL(1); }
    if (chunk) epilogue_chunk(p, P, tick);
    else epilogue_scan<DIR>(p, P, tick, pc);
    asm volatile("fence.proxy.async.global;" ::: "memory");
    epi_group_sync();
    if (threadIdx.x == 64) { STAMP(S, bar, 1); TL(8); grid_arrive(ctr); }
  }
  ++bar;
}
// Shared-memory copies used by the persistent loops: the three EngineParams and, for the two scan tables, this CTA's
// job and epilogue context (job index = CTA index, every tick).  All threads call this; ends with a CTA barrier.
struct PersistShared {
  EngineParams P[3];
  PhaseCache pc[2];
};
__device__ __forceinline__ PersistShared* persist_shared_fill(uint8_t* area, const EngineParams* ph) {
  PersistShared* ps = reinterpret_cast<PersistShared*>(area);
  {
    const int* src = reinterpret_cast<const int*>(ph);
    int* dst = reinterpret_cast<int*>(ps->P);
    for (int i = threadIdx.x; i < (int)(3 * sizeof(EngineParams) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const int rank = (int)blockIdx.x - ph[0].cta0;
  for (int k = 0; k < 2; ++k)
    if (rank >= 0 && rank < ph[k].njobs) {
      const int* src = reinterpret_cast<const int*>(ph[k].jobs + rank);
      int* dst = reinterpret_cast<int*>(&ps->pc[k].job);
      for (int i = threadIdx.x; i < (int)(sizeof(Job) / 4); i += blockDim.x) dst[i] = src[i];
    }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < 2; ++k)
      if (rank >= 0 && rank < ph[k].njobs) {
        ps->pc[k].epi = make_epi_local(ps->pc[k].job, ph[k].ctx);
        ps->pc[k].valid = 1;
      }
  __syncthreads();
  return ps;
}
static_assert(sizeof(PersistShared) <= 1792, "PersistShared must fit behind the pipeline barriers");

// =========================================================================
// Grouped persistent scans.  The 148 CTAs are partitioned into three LAYER GROUPS, each with its own barrier counter
// and its own tick loop: group l runs the recurrence of layer l + 1 (gates phase, candidate phase; group 0 also the
// two attention stages) and, once per chunk of Tc steps, the hoisted "chunk" products that feed it.  The groups are
// coupled only through chunk-level dependencies (forward: group l waits until group l - 1 has finished the steps of
// the chunk; backward: group l waits for group l + 1), checked against the other group's monotonic barrier counter.
// A tick therefore costs max over groups instead of the sum of all phases of all layers: layers 2 / 3 (two phases per
// step) and the throughput-bound chunk products hide under the four dependent phases of layer 1 + attention.
// =========================================================================
// accumulators of the grouped kernels: two buffers GROUP_ACC_STRIDE columns apart at the bottom of the CTA's tensor
// memory; the columns above hold the resident weight tiles of the CTA's two scan jobs (Job::res_col / res_kb, set by
// the host: api.cu assign_resident)
constexpr int GROUP_ACC_STRIDE = 128;
constexpr int GROUP_RES_COL0 = 2 * GROUP_ACC_STRIDE;
constexpr int GROUP_RES_KB = (512 - GROUP_RES_COL0) / 64;   // resident k blocks (hi + lo plane, 64 columns) per CTA
__device__ __forceinline__ void group_resident_setup(Pipe& p, const PersistShared* ps) {
  p.acc_stride = GROUP_ACC_STRIDE;
  for (int k = 0; k < 2; ++k) resident_preload(p, ps->P[k], &ps->pc[k]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
struct GroupSched {
  EngineParams ph[3];   // [0] / [1] the two recurrent phases of the layer, [2] its chunk table (njobs may be 0)
  int cta0, ncta;
};
constexpr int BAR_STRIDE = 32;   // uints between the barrier counters of the groups (128 bytes)
struct ScanFwdGParams {
  GroupSched g[3];
  AttnFwdArgs att;     // pointers of step 0
  long long s_h1, s_k, s_w, s_wp, s_phi, s_ab, s_e;   // per-step strides (elements)
  int T, Tc;
  int att_parts, att_slices;
  unsigned int* bars;           // [3][BAR_STRIDE]
  unsigned long long* stamps;   // debug: [cta][bar][2]
  int stamp_bars;
  int prefetch;
};
struct ScanBwdGParams {
  GroupSched g[3];
  AttnBwdArgs att;     // pointers of step 0
  long long s_dw, s_ab, s_e, s_k, s_dh1, s_datt, s_dattp;
  const ScanCtx* ctx;
  int T, Tc;
  unsigned int* bars;
  unsigned long long* stamps;
  int stamp_bars;
  int prefetch;
  int fused_pre;   // groups 1 / 2 run the GRU pre-pass inside the state-dgrad finish (QF_FUSED_PRE set in their tables)
  unsigned long long* tl_buf;   // debug: attention-backward milestones of batch row 0 at step tl_tick go to
  int tl_tick;                  // tl_buf[2 * 148 * 16 ..] (tools/group_timeline.py)
};
// barriers group `g` of the forward kernel has completed once it has finished `nt` steps
__device__ __forceinline__ unsigned int fwd_bars_after(int g, int nt, int Tc, bool has_chunk) {
  const unsigned int chunks = has_chunk ? (unsigned int)((nt + Tc - 1) / Tc) : 0u;
  return (g == 0 ? 4u : 2u) * (unsigned int)nt + chunks;
}
// backward kernel: group 0 runs three barriers per step (attention backward + pre-pass, two dgrad phases); groups
// 1 / 2 run the stand-alone pre-pass only at their first step (afterwards it is fused into the state-dgrad finish,
// engine.cuh QF_FUSED_PRE) and one chunk event per finished range
__device__ __forceinline__ unsigned int bwd_bars_after(int g, int nt, int Tc, bool has_chunk, bool fused) {
  const unsigned int chunks = has_chunk ? (unsigned int)((nt + Tc - 1) / Tc) : 0u;   // (the last range may be short)
  return ((g == 0 || !fused) ? 3u * (unsigned int)nt : 2u * (unsigned int)nt + (nt > 0 ? 1u : 0u)) + chunks;
}

__global__ void __launch_bounds__(ENGINE_THREADS, 1) scan_fwd_grouped(const ScanFwdGParams S) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Pipe p;
  int max_cols = S.g[0].ph[0].n_cols;
  for (int g = 1; g < 3; ++g)
    if (S.g[g].ph[2].njobs > 0 && S.g[g].ph[2].n_cols > max_cols) max_cols = S.g[g].ph[2].n_cols;
  float* att_sh = reinterpret_cast<float*>(pipe_setup(p, align_smem(smem_raw), max_cols));
  p.stg = reinterpret_cast<uint8_t*>(att_sh); p.stg_bytes = ATT_SMEM_BYTES;   // finish operands (GEMM phases only)
  const int gi = (int)blockIdx.x >= S.g[2].cta0 ? 2 : ((int)blockIdx.x >= S.g[1].cta0 ? 1 : 0);
  const GroupSched& G = S.g[gi];
  const PersistShared* ps = persist_shared_fill(p.cache_area, G.ph);
  group_resident_setup(p, ps);
  const int rank = (int)blockIdx.x - G.cta0, ncta = G.ncta;
  unsigned int* ctr = S.bars + gi * BAR_STRIDE;
  const unsigned int* lower = gi > 0 ? S.bars + (gi - 1) * BAR_STRIDE : nullptr;
  const int lower_ncta = gi > 0 ? S.g[gi - 1].ncta : 0;
  const bool lower_chunk = gi > 0 && S.g[gi - 1].ph[2].njobs > 0;
  const bool has_chunk = ps->P[2].njobs > 0;
  unsigned int bar = 0;
  for (int t = 0; t < S.T; ++t) {
    if (has_chunk && t % S.Tc == 0) {
      // hoisted products of chunk e = t / Tc: their operands are the lower layers' states of steps [t, t + Tc)
      const int need = min(t + S.Tc, S.T);
      const unsigned int xt = fwd_bars_after(gi - 1, need, S.Tc, lower_chunk) * (unsigned int)lower_ncta;
      group_gemm_phase<1>(p, ps->P[2], t / S.Tc, S, ctr, ncta, bar, nullptr, true, lower, xt);
    }
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) group_gemm_phase<1>(p, ps->P[ph], t, S, ctr, ncta, bar, &ps->pc[ph], false);
    if (gi == 0) {
      AttnFwdArgs a = S.att;
      a.h1 += t * S.s_h1; a.k_prev += t * S.s_k; a.k_out += t * S.s_k; a.w_out += t * S.s_w;
      a.w_hi += t * S.s_wp; a.w_lo += t * S.s_wp; a.phi_out += t * S.s_phi; a.ab_out += t * S.s_ab;
      a.e_out += t * S.s_e;
      // stage 1 (att_slices CTAs): K-sliced partial projections of h1_t ; stage 2 (att_parts CTAs per batch row):
      // window + context slice
      if (threadIdx.x == 0) { grid_wait(ctr, bar * ncta); STAMP(S, bar, 0); }
      __syncthreads();
      for (int sl = rank; sl < S.att_slices; sl += ncta) attention_proj_slice(a, sl, att_sh);
      __syncthreads();
      if (threadIdx.x == 32) { STAMP(S, bar, 1); grid_arrive(ctr); }   // (not thread 0: it is the next phase's TMA producer)
      ++bar;
      const int nwork = a.B * S.att_parts;
      if (threadIdx.x == 0) { grid_wait(ctr, bar * ncta); STAMP(S, bar, 0); }
      __syncthreads();
      for (int i = rank; i < nwork; i += ncta)
        attention_window_part<true>(a, i / S.att_parts, i % S.att_parts, S.att_parts, S.att_slices, att_sh);
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 32) { STAMP(S, bar, 1); grid_arrive(ctr); }   // (not thread 0: it is the next phase's TMA producer)
      ++bar;
    }
  }
  pipe_teardown(p);
}

__global__ void __launch_bounds__(ENGINE_THREADS, 1) scan_bwd_grouped(const ScanBwdGParams S) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Pipe p;
  int max_cols = S.g[0].ph[0].n_cols;
  for (int g = 0; g < 3; ++g)
    if (S.g[g].ph[2].njobs > 0 && S.g[g].ph[2].n_cols > max_cols) max_cols = S.g[g].ph[2].n_cols;
  float* att_sh = reinterpret_cast<float*>(pipe_setup(p, align_smem(smem_raw), max_cols));
  p.stg = reinterpret_cast<uint8_t*>(att_sh); p.stg_bytes = ATT_SMEM_BYTES;
  const int gi = (int)blockIdx.x >= S.g[2].cta0 ? 2 : ((int)blockIdx.x >= S.g[1].cta0 ? 1 : 0);
  const GroupSched& G = S.g[gi];
  const PersistShared* ps = persist_shared_fill(p.cache_area, G.ph);
  group_resident_setup(p, ps);
  const int rank = (int)blockIdx.x - G.cta0, ncta = G.ncta;
  unsigned int* ctr = S.bars + gi * BAR_STRIDE;
  const unsigned int* upper = gi < 2 ? S.bars + (gi + 1) * BAR_STRIDE : nullptr;
  const int upper_ncta = gi < 2 ? S.g[gi + 1].ncta : 0;
  const bool upper_chunk = gi < 2 && S.g[gi + 1].ph[2].njobs > 0;
  const bool has_chunk = ps->P[2].njobs > 0;
  const ScanCtx& c = *S.ctx;
  unsigned int bar = 0;
  for (int k = 0; k < S.T; ++k) {
    const int s = S.T - 1 - k;
    // What the step needs from the upper group (chunk dgrads accumulated into this layer's dh / dw slots):
    // (a) first step of a range: the upper group's chunk dgrads of this range (slots s + 1 ..) ;
    // (b) last step of a range: its chunk dgrads of the NEXT range, whose highest slot is the slot s this step
    //     read-modify-writes (the only slot two ranges share) -- and which the fused pre-pass of step s - 1 reads.
    const unsigned int* xw = nullptr;
    unsigned int xt = 0;
    //     With the fused pre-pass (groups 1 / 2) that read-modify-write happens one step EARLIER -- in the state-dgrad
    //     finish of the step before the last one of the range -- so the wait moves one step up as well.
    const int ahead = (gi > 0 && S.fused_pre) ? 2 : 1;
    if (upper && k % S.Tc == 0) {
      xw = upper; xt = bwd_bars_after(gi + 1, min(k + S.Tc, S.T), S.Tc, upper_chunk, S.fused_pre != 0) * (unsigned int)upper_ncta;
    } else if (upper && (k + ahead) % S.Tc == 0 && k + ahead < S.T) {
      xw = upper; xt = bwd_bars_after(gi + 1, min(k + ahead + S.Tc, S.T), S.Tc, upper_chunk, S.fused_pre != 0) * (unsigned int)upper_ncta;
    }
    // phase 0: everything of the step that is elementwise in dh_s -- group 0: attention backward of step s, each row
    // followed by its GRU pre-pass of layer 1 ; groups 1 / 2: the GRU pre-pass of their layer, needed as a phase of
    // its own only at the first step (later steps: fused into the previous step's state-dgrad finish)
    if (gi == 0 || k == 0 || !S.fused_pre) {
      if (threadIdx.x == 0) {
        if (bar) grid_wait(ctr, bar * ncta);
        if (xw) grid_wait(xw, xt);
        STAMP(S, bar, 0);
      }
      __syncthreads();
      if (gi == 0) {
        AttnBwdArgs a = S.att;
        a.dbg = (S.tl_buf && k == S.tl_tick) ? S.tl_buf + 2 * 148 * 16 : nullptr;
        a.dw += s * S.s_dw; a.ab += s * S.s_ab; a.e += s * S.s_e; a.kappa += s * S.s_k; a.dh1 += s * S.s_dh1;
        a.datt += s * S.s_datt; a.datt_hi += s * S.s_dattp; a.datt_lo += s * S.s_dattp;
        for (int b = rank; b < a.B; b += ncta) attention_bwd_body(a, b, att_sh, &c, s);
      } else {
        gru_bwd_pre_rows(c, gi, s, 0, c.B, rank * (int)blockDim.x + (int)threadIdx.x, ncta * (int)blockDim.x);
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 32) { STAMP(S, bar, 1); grid_arrive(ctr); }   // (not thread 0: it is the next phase's TMA producer)
      ++bar;
      xw = nullptr;
    }
    group_gemm_phase<2>(p, ps->P[0], k, S, ctr, ncta, bar, &ps->pc[0], false, xw, xt);
    group_gemm_phase<2>(p, ps->P[1], k, S, ctr, ncta, bar, &ps->pc[1], false);
    if (has_chunk && ((k + 1) % S.Tc == 0 || k == S.T - 1))
      group_gemm_phase<2>(p, ps->P[2], k / S.Tc, S, ctr, ncta, bar, nullptr, true);
  }
  pipe_teardown(p);
}

// =========================================================================
// Operand packing: fp32 -> bf16 hi/lo planes, optionally transposed.
// =========================================================================
// dst planes [R_out][ld] ; src fp32 [rows][cols] row-major.
// transpose=1: dst[c][r] = src[r][c] (dst rows = src cols)
// tiled_nkb > 0: the destination is tile-contiguous: element (r, c) of the (possibly transposed) result goes to
// ((r/128)*tiled_nkb + c/64)*128*64 + (r%128)*64 + c%64
__device__ __forceinline__ long long pack_dst_index(long long r, long long c, long long dst_ld, int tiled_nkb) {
  if (tiled_nkb <= 0) return r * dst_ld + c;
  return (((r >> 7) * tiled_nkb + (c >> 6)) * 128 + (r & 127)) * 64 + (c & 63);
}
__global__ void pack_planes_kernel(const float* __restrict__ src, long long src_ld, int rows, int cols,
                                   bf16* __restrict__ hi, bf16* __restrict__ lo, long long dst_ld,
                                   int transpose, int tiled_nkb) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (!transpose) {
    for (int j = ty; j < 32; j += 8) {
      const int r = by + j, c = bx + tx;
      if (r < rows && c < cols) {
        bf16 hh, ll;
        split_bf16(src[(long long)r * src_ld + c], hh, ll);
        const long long o = pack_dst_index(r, c, dst_ld, tiled_nkb);
        hi[o] = hh;
        lo[o] = ll;
      }
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      const int r = by + j, c = bx + tx;
      tile[j][tx] = (r < rows && c < cols) ? src[(long long)r * src_ld + c] : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = bx + j, r = by + tx;  // dst row = c, dst col = r
      if (r < rows && c < cols) {
        bf16 hh, ll;
        split_bf16(tile[tx][j], hh, ll);
        const long long o = pack_dst_index(c, r, dst_ld, tiled_nkb);
        hi[o] = hh;
        lo[o] = ll;
      }
    }
  }
}

// bf16 plane transpose: dst[c][r] = src[r][c]   (for the weight-gradient operands)
// 64 x 64 tiles, 4-byte (bf16x2) global accesses on both sides.  Requires even leading dimensions.
__global__ void transpose_plane64_kernel(const bf16* __restrict__ src, long long src_ld, long long rows, int cols,
                                         bf16* __restrict__ dst, long long dst_ld) {
  __shared__ bf16 tile[64][66];
  const long long by = (long long)blockIdx.y * 64;
  const int bx = blockIdx.x * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int j = ty; j < 64; j += 8) {
    const long long r = by + j;
    const int c = bx + 2 * tx;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.0f, 0.0f);
    if (r < rows && c + 1 < cols) v = *reinterpret_cast<const __nv_bfloat162*>(src + r * src_ld + c);
    else if (r < rows && c < cols) v.x = src[r * src_ld + c];
    tile[j][2 * tx] = v.x;
    tile[j][2 * tx + 1] = v.y;
  }
  __syncthreads();
  for (int j = ty; j < 64; j += 8) {
    const int c = bx + j;              // dst row
    const long long r = by + 2 * tx;   // dst column pair
    if (c < cols) {
      if (r + 1 < rows) {
        __nv_bfloat162 v;
        v.x = tile[2 * tx][j];
        v.y = tile[2 * tx + 1][j];
        *reinterpret_cast<__nv_bfloat162*>(dst + (long long)c * dst_ld + r) = v;
      } else if (r < rows) {
        dst[(long long)c * dst_ld + r] = tile[2 * tx][j];
      }
    }
  }
}

__global__ void transpose_plane_kernel(const bf16* __restrict__ src, long long src_ld, long long rows, int cols,
                                       bf16* __restrict__ dst, long long dst_ld) {
  __shared__ bf16 tile[32][34];
  const long long by = (long long)blockIdx.y * 32;
  const int bx = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int j = ty; j < 32; j += 8) {
    const long long r = by + j;
    const int c = bx + tx;
    tile[j][tx] = (r < rows && c < cols) ? src[r * src_ld + c] : __float2bfloat16(0.0f);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = bx + j;
    const long long r = by + tx;
    if (r < rows && c < cols) dst[(long long)c * dst_ld + r] = tile[tx][j];
  }
}

// fp32 [T][B][F] (optionally + level * noise) -> planes [T][Np][Fp]
__global__ void frames_to_planes_kernel(const float* __restrict__ src, const float* __restrict__ noise,
                                        float level, int T, int B, int F, int Np, int Fp,
                                        bf16* __restrict__ hi, bf16* __restrict__ lo, float* __restrict__ copy) {
  const long long n = (long long)T * B * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long tb = i / F;
    const int b = (int)(tb % B);
    const long long t = tb / B;
    float x = src[i];
    if (noise) x += level * noise[i];
    if (copy) copy[i] = x;
    bf16 hh, ll;
    split_bf16(x, hh, ll);
    const long long o = (t * Np + b) * Fp + f;
    hi[o] = hh;
    lo[o] = ll;
  }
}

// state rows [B][F] fp32 -> fp32 slot + planes slot
__global__ void state_to_slot_kernel(const float* __restrict__ src, long long src_bstride, int B, int F, int Np,
                                     int Fp, float* __restrict__ dst, bf16* __restrict__ hi,
                                     bf16* __restrict__ lo) {
  const long long n = (long long)B * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / F), f = (int)(i % F);
    const float x = src[(long long)b * src_bstride + f];
    if (dst) dst[i] = x;
    if (hi) {
      bf16 hh, ll;
      split_bf16(x, hh, ll);
      hi[(long long)b * Fp + f] = hh;
      lo[(long long)b * Fp + f] = ll;
    }
  }
}

// =========================================================================
// Strided SIMT GEMM for the tiny / odd products:
//   C[m][n] = alpha * sum_k A(m,k) * B(k,n) + beta * C[m][n] (+ bias[n])
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn], C row-major ldc.
// Optional row gather on A: m -> a_rows[m].
// =========================================================================
struct SGemm {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  float* C; long long ldc;
  const float* bias;
  const int* a_gather;
  int M, N, K;
  float alpha, beta;
  int k_chunk;            // split-K: block z reduces k in [z*k_chunk, (z+1)*k_chunk) into C + z*c_zstride
  long long c_zstride;
};
__global__ void __launch_bounds__(256) sgemm_kernel(const SGemm g) {
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int kb = blockIdx.z * g.k_chunk, ke = min(g.K, kb + g.k_chunk);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = kb; k0 < ke; k0 += 32) {
    for (int j = ty; j < 32; j += 8) {
      const int m = m0 + j, k = k0 + tx;
      float x = 0.0f;
      if (m < g.M && k < ke) {
        const long long mr = g.a_gather ? g.a_gather[m] : m;
        x = g.A[mr * g.sam + (long long)k * g.sak];
      }
      As[j][tx] = x;
      const int kk = k0 + j, n = n0 + tx;
      Bs[j][tx] = (kk < ke && n < g.N) ? g.B[(long long)kk * g.sbk + (long long)n * g.sbn] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float bv = Bs[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + 8 * i][k], bv, acc[i]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < g.N) {
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 8 * i;
      if (m < g.M) {
        float y = g.alpha * acc[i];
        if (g.bias) y += g.bias[n];
        float* p = g.C + (long long)blockIdx.z * g.c_zstride + (long long)m * g.ldc + n;
        if (g.beta != 0.0f) y += g.beta * *p;
        *p = y;
      }
    }
  }
}

// column sums: out[f] (+)= sum_rows src[r][f]      (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ src, long long ld, long long rows, int cols,
                              float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.0f;
  if (c < cols)
    for (long long r = ty; r < rows; r += 8) s += src[r * ld + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.0f;
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// Two-stage, deterministic column sums for tall matrices: stage 1 sums a chunk of rows per block into
// partial[chunk][col]; stage 2 adds the chunks in chunk order.
__global__ void colsum_partial_kernel(const float* __restrict__ src, long long ld, long long rows, int cols,
                                      long long rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float s = 0.0f;
  if (c < cols)
    for (long long r = r0 + ty; r < r1; r += 8) s += src[r * ld + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.0f;
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    partial[(long long)blockIdx.y * cols + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int chunks, int cols, float* __restrict__ out,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float t = 0.0f;
  for (int k = 0; k < chunks; ++k) t += partial[(long long)k * cols + c];
  out[c] = accumulate ? out[c] + t : t;
}

__global__ void add_vec_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

// dst[t][b][f] = sum over t of src  -> [b][f]   (speaker-path gradients)
__global__ void timesum_kernel(const float* __restrict__ src, int T, long long bf, float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < bf;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int t = 0; t < T; ++t) s += src[(long long)t * bf + i];
    dst[i] = s;
  }
}

// scatter-add rows: dst[idx[m]][:] += src[m][:]   (lookup-table gradients).  Deterministic: block (f-chunk, r)
// scans all m, warp w taking m = w, w + nwarp, ...; the per-warp sums are added in warp order.
__global__ void __launch_bounds__(1024) scatter_rows_kernel(const float* __restrict__ src, long long src_ld,
                                                            const int* __restrict__ idx, int M, int F, int rows,
                                                            float* __restrict__ dst, long long dst_ld) {
  __shared__ float part[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int r = blockIdx.y, f = blockIdx.x * 32 + lane;
  float s = 0.0f;
  if (f < F)
    for (int m = warp; m < M; m += nwarp)
      if (__ldg(idx + m) == r) s += src[(long long)m * src_ld + f];
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && f < F && r < rows) {
    float t = 0.0f;
    for (int w = 0; w < nwarp; ++w) t += part[w][lane];
    dst[(long long)r * dst_ld + f] += t;
  }
}

// dst[c][r] = src[r][c]  (fp32)
__global__ void transpose_f32_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  const long long n = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(long long)c * rows + r] = src[i];
  }
}

// =========================================================================
// _simple_norm (model.py:24-27): (x - mean) / (1e-5 + std) over a feature range, population std, no affine.
// Used only with layer_norm=True.  One CTA per (row, part); a row may hold two independently normalised
// parts (the [cell | gates] outputs of one Fork pair).
// =========================================================================
struct NormParts { int off[2]; int n[2]; int poff[2]; int count; };

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// y[r][off + j] (= | +=) norm(x[r][off .. off + n))[j]
__global__ void __launch_bounds__(256) rownorm_fwd_kernel(const float* __restrict__ x, long long xld,
                                                          float* __restrict__ y, long long yld, NormParts P,
                                                          int accumulate) {
  __shared__ float red[8];
  const long long r = blockIdx.x;
  const int part = blockIdx.y;
  const int off = P.off[part], n = P.n[part];
  const float* xr = x + r * xld + off;
  float* yr = y + r * yld + off;
  float s = 0.0f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) s += xr[j];
  const float mean = block_sum(s, red) / n;
  float v = 0.0f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) { const float d = xr[j] - mean; v += d * d; }
  const float sd = sqrtf(block_sum(v, red) / n);
  const float inv = 1.0f / (1e-5f + sd);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float o = (xr[j] - mean) * inv;
    yr[j] = accumulate ? yr[j] + o : o;
  }
}

// dx = d/dx of norm(x) applied to dy ; optional bf16 planes of dx at plane row (r / Bv) * Np + r % Bv, column
// poff[part] + j
__global__ void __launch_bounds__(256) rownorm_bwd_kernel(const float* __restrict__ dy, long long dyld,
                                                          const float* __restrict__ x, long long xld,
                                                          float* __restrict__ dx, long long dxld, bf16* hi,
                                                          bf16* lo, long long pld, int Bv, int Np, NormParts P) {
  __shared__ float red[8];
  const long long r = blockIdx.x;
  const int part = blockIdx.y;
  const int off = P.off[part], n = P.n[part];
  const float* xr = x + r * xld + off;
  const float* dyr = dy + r * dyld + off;
  float s = 0.0f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) s += xr[j];
  const float mean = block_sum(s, red) / n;
  float v = 0.0f, sdy = 0.0f, dot = 0.0f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float d = xr[j] - mean;
    v += d * d;
    sdy += dyr[j];
    dot += dyr[j] * d;
  }
  const float sd = sqrtf(block_sum(v, red) / n);
  const float mdy = block_sum(sdy, red) / n;
  const float dotv = block_sum(dot, red);
  const float sv = 1e-5f + sd;
  const float k2 = (sd > 0.0f) ? dotv / ((float)n * sd * sv * sv) : 0.0f;
  const long long prow = (r / Bv) * Np + (r % Bv);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float g = (dyr[j] - mdy) / sv - (xr[j] - mean) * k2;
    dx[r * dxld + off + j] = g;
    if (hi) {
      bf16 hh, ll;
      split_bf16(g, hh, ll);
      hi[prow * pld + P.poff[part] + j] = hh;
      lo[prow * pld + P.poff[part] + j] = ll;
    }
  }
}

// dst[i] = sum_j src_j[i]  (null sources ignored)
__global__ void vec_sum_kernel(float* __restrict__ dst, int n, const float* s0, const float* s1, const float* s2,
                               const float* s3) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (s0) v += s0[i];
    if (s1) v += s1[i];
    if (s2) v += s2[i];
    if (s3) v += s3[i];
    dst[i] = v;
  }
}

// dst[t][b][f] = src[b][f] for all t
__global__ void bcast_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int T, long long bf) {
  const long long n = (long long)T * bf;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i % bf];
}

__global__ void mul_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b,
                           long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = a[i] * b[i];
}

// gather the per-character encoder projections: proj [NC][2][3E] -> xi[dir] [L][N][E], xg[dir] [L][N][2E]
// label of (l, n): time_axis 0 -> labels[l*U + n] (L=B, N=U) ; 1 -> labels[n*U + l] (L=U, N=B)
__global__ void enc_gather_kernel(const float* __restrict__ proj, const int* __restrict__ labels, int L, int N,
                                  int U, int time_axis, int E, float* xi0, float* xg0, float* xi1, float* xg1,
                                  int* __restrict__ lab_out) {
  const long long n = (long long)L * N * 6 * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % (6 * E));
    const long long ln = i / (6 * E);
    const int l = (int)(ln / N), nn = (int)(ln % N);
    const int ch = time_axis == 0 ? labels[(long long)l * U + nn] : labels[(long long)nn * U + l];
    if (j == 0) lab_out[ln] = ch;
    const float v = proj[(long long)ch * 6 * E + j];
    const int dir = j / (3 * E), jj = j % (3 * E);
    float* xi = dir ? xi1 : xi0;
    float* xg = dir ? xg1 : xg0;
    if (jj < E) xi[ln * E + jj] = v;
    else xg[ln * 2 * E + (jj - E)] = v;
  }
}

// =========================================================================
// Emitter: MSE (model.py:757-764) and diagonal-GMM NLL (model.py:65-91, 774-781)
// =========================================================================
// One warp per frame.  pred layout for GMM: [mu (D*k) | sigma_hat (D*k) | coeff_hat (k)], index d*k + j.
struct EmitArgs {
  int N;          // frames = T*B
  int D, k, which;  // which: 0 MSE, 1 GMM
  int Dtot;       // row pitch of pred
  float eps;
  const float* pred;
  const float* target;  // [N][D]
  const float* mask;    // [N]
  float* cost_tb;       // [N]
  // backward
  float* dpred;         // [N][Dtot] fp32
  bf16* dpred_hi;       // planes [q*Np + r][Dp]
  bf16* dpred_lo;
  int B, Np, Dp;
  int DK, poff1, poff2;   // GMM: plane column of block f is poff_f + (idx - f*DK)
  const float* scale;   // device scalar: 1/(sum mask + 1e-5) or 1 (unnormalised)
};

__global__ void __launch_bounds__(256) emit_cost_kernel(const EmitArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.N) return;
  const float* p = a.pred + (long long)warp * a.Dtot;
  const float* y = a.target + (long long)warp * a.D;
  if (a.which == 0) {
    float s = 0.0f;
    for (int d = lane; d < a.D; d += 32) {
      const float e = p[d] - y[d];
      s += e * e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) a.cost_tb[warp] = s;
    return;
  }
  const int k = a.k, D = a.D;
  // lane j < k owns component j
  float inner = 0.0f, lw = -INFINITY;
  if (lane < k) {
#pragma unroll 7   // independent loads in flight (63 = 9 x 7 output dimensions)
    for (int d = 0; d < D; ++d) {
      const float mu = p[d * k + lane];
      const float sg = expf(p[D * k + d * k + lane]) + a.eps;
      const float df = y[d] - mu;
      inner += (df * df) / (sg * sg) + 2.0f * logf(sg) + 1.8378770664093453f;
    }
    inner *= -0.5f;
  }
  // softmax over coeff_hat
  float ch = (lane < k) ? p[2 * D * k + lane] : -INFINITY;
  float m = ch;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float ex = (lane < k) ? expf(ch - m) : 0.0f;
  float sum = ex;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float pi = ex / sum + a.eps;
  if (lane < k) lw = logf(pi) + inner;
  float mx = lw;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = (lane < k) ? expf(lw - mx) : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) a.cost_tb[warp] = -(logf(se) + mx);
}

// gradient of sum_n cost_tb[n] * mask[n] * scale wrt pred
__global__ void __launch_bounds__(256) emit_grad_kernel(const EmitArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.N) return;
  const float* p = a.pred + (long long)warp * a.Dtot;
  const float* y = a.target + (long long)warp * a.D;
  float* g = a.dpred + (long long)warp * a.Dtot;
  const int q = warp / a.B, r = warp % a.B;
  bf16* ghi = a.dpred_hi + ((long long)q * a.Np + r) * a.Dp;
  bf16* glo = a.dpred_lo + ((long long)q * a.Np + r) * a.Dp;
  const float sc = a.mask[warp] * a.scale[0];
  auto put = [&](int idx, float v) {
    g[idx] = v;
    bf16 hh, ll;
    split_bf16(v, hh, ll);
    int pc = idx;
    if (a.which == 1) pc = idx < a.DK ? idx : (idx < 2 * a.DK ? a.poff1 + (idx - a.DK) : a.poff2 + (idx - 2 * a.DK));
    ghi[pc] = hh;
    glo[pc] = ll;
  };
  if (a.which == 0) {
    for (int d = lane; d < a.D; d += 32) put(d, 2.0f * (p[d] - y[d]) * sc);
    return;
  }
  const int k = a.k, D = a.D;
  float inner = 0.0f, lw = -INFINITY;
  if (lane < k) {
#pragma unroll 7   // independent loads in flight (63 = 9 x 7 output dimensions)
    for (int d = 0; d < D; ++d) {
      const float mu = p[d * k + lane];
      const float sg = expf(p[D * k + d * k + lane]) + a.eps;
      const float df = y[d] - mu;
      inner += (df * df) / (sg * sg) + 2.0f * logf(sg) + 1.8378770664093453f;
    }
    inner *= -0.5f;
  }
  float ch = (lane < k) ? p[2 * D * k + lane] : -INFINITY;
  float m = ch;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float ex = (lane < k) ? expf(ch - m) : 0.0f;
  float sum = ex;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float sm = ex / sum;
  const float pi = sm + a.eps;
  if (lane < k) lw = logf(pi) + inner;
  float mx = lw;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = (lane < k) ? expf(lw - mx) : 0.0f;
  float sse = se;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
  const float rho = se / sse;                      // responsibility of component `lane`
  const float dpi = (lane < k) ? -(rho / pi) * sc : 0.0f;
  float dot = dpi * sm;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (lane < k) {
    put(2 * D * k + lane, sm * (dpi - dot));
    for (int d = 0; d < D; ++d) {
      const float mu = p[d * k + lane];
      const float esg = expf(p[D * k + d * k + lane]);
      const float sg = esg + a.eps;
      const float df = y[d] - mu;
      const float dmu = -(rho * df / (sg * sg)) * sc;
      const float dsg = -(rho * ((df * df) / (sg * sg * sg) - 1.0f / sg)) * sc;
      put(d * k + lane, dmu);
      put(D * k + d * k + lane, dsg * esg);
    }
  }
}

// masked mean: out[0] = sum(cost*mask)/(sum(mask)+1e-5) ; out[1] = sum(cost*mask); out[2] = sum(mask);
// out[3] = 1/(sum(mask)+1e-5).  Single block, fixed order => deterministic.
__global__ void __launch_bounds__(1024) masked_mean_kernel(const float* __restrict__ cost,
                                                            const float* __restrict__ mask, long long n,
                                                            float* __restrict__ out) {
  __shared__ float s1[32], s2[32];
  float a = 0.0f, b = 0.0f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    a += cost[i] * mask[i];
    b += mask[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.0f, y = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { x += s1[i]; y += s2[i]; }
    out[0] = x / (y + 1e-5f);
    out[1] = x;
    out[2] = y;
    out[3] = 1.0f / (y + 1e-5f);
  }
}

// =========================================================================
// GMM sampling, one step (model.py:94-118, 1024-1033).  One warp per row.
// pred: [B][Dtot] = [mu | sigma_hat | coeff_hat];  x = mu_j + sigma_j * normal.
// Noise is either injected (unis/normals non-null: parity tests) or drawn from a
// counter-based Philox-4x32-10 stream keyed by (seed, step, row).
// =========================================================================
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

struct SampleArgs {
  int B, D, k, Dtot, which;
  float eps, bias;           // sampling_bias
  const float* pred;         // [B][Dtot]
  const float* unis;         // [B] or null
  const float* normals;      // [B][D] or null
  unsigned long long seed; int step;
  const unsigned long long* seed_ptr;   // when set, the Philox key is read from here (graph replays with a new seed)
  float* x_out;              // [B][D]
  float* pi_out;             // [B][k] (GMM) / [B][D] (MSE: copy of x)
  bf16* x_hi; bf16* x_lo; int Np, Dp;   // planes of x for the feedback product (nullable)
};
__global__ void __launch_bounds__(128) sample_emit_kernel(const SampleArgs a_in) {
  SampleArgs a = a_in;
  if (a.seed_ptr) a.seed = *a.seed_ptr;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= a.B) return;
  const float* p = a.pred + (long long)row * a.Dtot;
  auto putx = [&](int d, float v) {
    a.x_out[(long long)row * a.D + d] = v;
    if (a.x_hi) {
      bf16 hh, ll;
      split_bf16(v, hh, ll);
      a.x_hi[(long long)row * a.Dp + d] = hh;
      a.x_lo[(long long)row * a.Dp + d] = ll;
    }
  };
  if (a.which == 0) {
    for (int d = lane; d < a.D; d += 32) {
      putx(d, p[d]);
      a.pi_out[(long long)row * a.D + d] = p[d];
    }
    return;
  }
  const int k = a.k, D = a.D;
  float ch = (lane < k) ? p[2 * D * k + lane] * (1.0f + a.bias) : -INFINITY;
  float m = ch;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float ex = (lane < k) ? expf(ch - m) : 0.0f;
  float sum = ex;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float pi = ex / sum + a.eps;
  if (lane < k) a.pi_out[(long long)row * k + lane] = pi;
  uint32_t rnd[4];
  float u;
  if (a.unis) u = a.unis[row];
  else {
    philox4x32((uint32_t)row, (uint32_t)a.step, 0u, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), rnd);
    u = u01(rnd[0]);
  }
  // first component whose running sum exceeds u (theano MultinomialFromUniform)
  int idx = 0;
  {
    float cum = 0.0f;
    bool done = false;
    for (int j = 0; j < k; ++j) {
      const float pj = __shfl_sync(0xffffffffu, pi, j);
      cum = __fadd_rn(cum, pj);
      if (!done && u < cum) { idx = j; done = true; }
    }
  }
  for (int d = lane; d < D; d += 32) {
    const float mu = p[d * k + idx];
    const float sg = expf(p[D * k + d * k + idx] - a.bias) + a.eps;
    float z;
    if (a.normals) z = a.normals[(long long)row * D + d];
    else {
      philox4x32((uint32_t)row, (uint32_t)a.step, (uint32_t)(d + 1), 1u, (uint32_t)a.seed,
                 (uint32_t)(a.seed >> 32), rnd);
      z = sqrtf(-2.0f * logf(u01(rnd[0]))) * cospif(2.0f * u01(rnd[1]));
    }
    putx(d, mu + sg * z);
  }
}

// =========================================================================
// Optimizer: StepClipping(threshold) + Adam (train.py:100-108), flat buffers.
// =========================================================================
__global__ void __launch_bounds__(1024) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                              double* __restrict__ partial) {
  __shared__ double sh[32];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)g[i];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
// Effective gradient = g * gscale * (mask_sum ? 1/(mask_sum[0] + 1e-5) : 1)   (data-parallel: the all-reduced
// gradients are un-normalised and the divisor is all-reduced with them).
// stats[0] = grad norm (of the effective gradient), stats[1] = clip multiplier, stats[2] = total multiplier
__global__ void clip_finalize_kernel(const double* __restrict__ partial, int nparts, float gscale,
                                     const float* __restrict__ mask_sum, float threshold,
                                     float* __restrict__ stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += partial[i];
    float sc = gscale;
    if (mask_sum) sc *= 1.0f / (mask_sum[0] + 1e-5f);
    const float norm = (float)sqrt(t) * fabsf(sc);
    stats[0] = norm;
    stats[1] = (norm < threshold) ? 1.0f : threshold / norm;
    stats[2] = stats[1] * sc;
  }
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const float* __restrict__ stats, float lr_t,
                            float b1, float b2, float eps) {
  const float mult = stats[2];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * mult;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// =========================================================================
// Encoder GRU recurrence (model.py:233-247 through blocks Bidirectional /
// GatedRecurrent).  xi/xg are the gathered Fork projections [L][N][E] / [L][N][2E].
// One CTA per (direction, chunk of ENC_ROWS rows); fp32 SIMT, weights from L2.
// Stashes z, r, c and the state sequence for the backward pass.
// =========================================================================
constexpr int ENC_ROWS = 8;
struct EncArgs {
  int L, N, E;
  const float* xi[2];     // [L][N][E]
  const float* xg[2];     // [L][N][2E]
  const float* Wg[2];     // [E][2E] state_to_gates
  const float* Ws[2];     // [E][E]  state_to_state
  const float* s0[2];     // [E] initial_state
  float* out;             // [L][N][2E]  forward | backward halves
  float* z[2]; float* r[2]; float* c[2];   // [L][N][E]
  float* sprev[2];        // [L][N][E] state entering step i
  // backward
  const float* dout;      // [L][N][2E]
  float* dxi[2]; float* dxg[2];
  float* ds0[2];          // [N][E] gradient wrt the initial state rows (summed later)
  int w_in_smem;          // stage the recurrent weights in shared memory (they fit: 3 E^2 floats)
};

__global__ void __launch_bounds__(256) encoder_fwd_kernel(const EncArgs a) {
  extern __shared__ float sh[];
  const int E = a.E, dir = blockIdx.y;
  const int n0 = blockIdx.x * ENC_ROWS;
  const int rows = min(ENC_ROWS, a.N - n0);
  float* s = sh;                    // [ENC_ROWS][E]
  float* rs = s + ENC_ROWS * E;     // [ENC_ROWS][E]
  float* g = rs + ENC_ROWS * E;     // [ENC_ROWS][2E]
  const int tid = threadIdx.x;
  // recurrent weights staged in shared memory once (192 KB at E = 128) when they fit
  const float* Wg = a.Wg[dir];
  const float* Ws = a.Ws[dir];
  if (a.w_in_smem) {
    float* wg_s = g + ENC_ROWS * 2 * E;
    float* ws_s = wg_s + 2 * E * E;
    for (int i = tid; i < 2 * E * E; i += blockDim.x) wg_s[i] = Wg[i];
    for (int i = tid; i < E * E; i += blockDim.x) ws_s[i] = Ws[i];
    Wg = wg_s; Ws = ws_s;
  }
  for (int i = tid; i < ENC_ROWS * E; i += blockDim.x) s[i] = a.s0[dir][i % E];
  __syncthreads();
  for (int step = 0; step < a.L; ++step) {
    const int i = dir == 0 ? step : a.L - 1 - step;
    // gates.  Register-tiled: four k per pass, the state rows read as broadcast 128-bit values -- 12 shared-memory
    // loads per 32 FMAs instead of 36 (the loop was shared-memory-issue bound: 9 us per encoder step)
    for (int j = tid; j < 2 * E; j += blockDim.x) {
      float acc[ENC_ROWS];
#pragma unroll
      for (int n = 0; n < ENC_ROWS; ++n) acc[n] = 0.0f;
      if ((E & 3) == 0) {
        for (int k = 0; k < E; k += 4) {
          const float w0 = Wg[(long long)k * 2 * E + j], w1 = Wg[(long long)(k + 1) * 2 * E + j];
          const float w2 = Wg[(long long)(k + 2) * 2 * E + j], w3 = Wg[(long long)(k + 3) * 2 * E + j];
#pragma unroll
          for (int n = 0; n < ENC_ROWS; ++n) {
            const float4 sv = *reinterpret_cast<const float4*>(s + n * E + k);
            acc[n] = fmaf(sv.x, w0, acc[n]); acc[n] = fmaf(sv.y, w1, acc[n]);
            acc[n] = fmaf(sv.z, w2, acc[n]); acc[n] = fmaf(sv.w, w3, acc[n]);
          }
        }
      } else
      for (int k = 0; k < E; ++k) {
        const float w = Wg[(long long)k * 2 * E + j];
#pragma unroll
        for (int n = 0; n < ENC_ROWS; ++n) acc[n] = fmaf(s[n * E + k], w, acc[n]);
      }
      for (int n = 0; n < rows; ++n) {
        const long long o = ((long long)i * a.N + n0 + n);
        g[n * 2 * E + j] = sigmoidf_exact(acc[n] + a.xg[dir][o * 2 * E + j]);
      }
    }
    __syncthreads();
    for (int e = tid; e < rows * E; e += blockDim.x) {
      const int n = e / E, f = e % E;
      const long long o = ((long long)i * a.N + n0 + n) * E + f;
      const float z = g[n * 2 * E + f], r = g[n * 2 * E + E + f];
      a.z[dir][o] = z;
      a.r[dir][o] = r;
      a.sprev[dir][o] = s[n * E + f];
      rs[n * E + f] = s[n * E + f] * r;
    }
    __syncthreads();
    for (int j = tid; j < E; j += blockDim.x) {
      float acc[ENC_ROWS];
#pragma unroll
      for (int n = 0; n < ENC_ROWS; ++n) acc[n] = 0.0f;
      if ((E & 3) == 0) {
        for (int k = 0; k < E; k += 4) {
          const float w0 = Ws[(long long)k * E + j], w1 = Ws[(long long)(k + 1) * E + j];
          const float w2 = Ws[(long long)(k + 2) * E + j], w3 = Ws[(long long)(k + 3) * E + j];
#pragma unroll
          for (int n = 0; n < ENC_ROWS; ++n) {
            const float4 sv = *reinterpret_cast<const float4*>(rs + n * E + k);
            acc[n] = fmaf(sv.x, w0, acc[n]); acc[n] = fmaf(sv.y, w1, acc[n]);
            acc[n] = fmaf(sv.z, w2, acc[n]); acc[n] = fmaf(sv.w, w3, acc[n]);
          }
        }
      } else
      for (int k = 0; k < E; ++k) {
        const float w = Ws[(long long)k * E + j];
#pragma unroll
        for (int n = 0; n < ENC_ROWS; ++n) acc[n] = fmaf(rs[n * E + k], w, acc[n]);
      }
      for (int n = 0; n < rows; ++n) {
        const long long o = ((long long)i * a.N + n0 + n);
        const float cc = tanhf(acc[n] + a.xi[dir][o * E + j]);
        const float z = g[n * 2 * E + j];
        const float sn = cc * z + s[n * E + j] * (1.0f - z);
        a.c[dir][o * E + j] = cc;
        a.out[o * 2 * E + dir * E + j] = sn;
        g[n * 2 * E + E + j] = sn;  // park the new state in the (now dead) reset half
      }
    }
    __syncthreads();
    for (int e = tid; e < rows * E; e += blockDim.x) {
      const int n = e / E, f = e % E;
      s[n * E + f] = g[n * 2 * E + E + f];
    }
    __syncthreads();
  }
}

// reverse-time encoder recurrence: consumes dout, writes dxi / dxg and ds0.
// Needs W^T products: d(rs) = da_c * Ws^T, ds += da_g * Wg^T  (uncoalesced reads of W rows are fine here:
// 2 x 196 KB of weights stay in L1/L2).
__global__ void __launch_bounds__(256) encoder_bwd_kernel(const EncArgs a) {
  extern __shared__ float sh[];
  const int E = a.E, dir = blockIdx.y;
  const int n0 = blockIdx.x * ENC_ROWS;
  const int rows = min(ENC_ROWS, a.N - n0);
  float* ds = sh;                      // [ENC_ROWS][E]
  float* dac = ds + ENC_ROWS * E;      // [ENC_ROWS][E]
  float* dag = dac + ENC_ROWS * E;     // [ENC_ROWS][2E]
  float* dsn = dag + ENC_ROWS * 2 * E; // [ENC_ROWS][E] next ds
  const int tid = threadIdx.x;
  // transposed weights in shared memory: WsT[j][k] = Ws[k][j], WgT[j][k] = Wg[k][j]  (k fastest: the threads of
  // a warp own consecutive k, so the reads below are conflict free; from global they were 32-way uncoalesced)
  float* wsT = dsn + ENC_ROWS * E;
  float* wgT = wsT + E * E;
  if (a.w_in_smem) {
    for (int i = tid; i < E * E; i += blockDim.x) { const int k = i / E, j = i % E; wsT[j * E + k] = a.Ws[dir][i]; }
    for (int i = tid; i < 2 * E * E; i += blockDim.x) { const int k = i / (2 * E), j = i % (2 * E); wgT[j * E + k] = a.Wg[dir][i]; }
  }
  for (int i = tid; i < ENC_ROWS * E; i += blockDim.x) ds[i] = 0.0f;
  __syncthreads();
  for (int step = a.L - 1; step >= 0; --step) {
    const int i = dir == 0 ? step : a.L - 1 - step;
    for (int e = tid; e < ENC_ROWS * E; e += blockDim.x) {
      const int n = e / E, f = e % E;
      float v_dac = 0.0f, v_dz = 0.0f, v_keep = 0.0f;
      if (n < rows) {
        const long long o = ((long long)i * a.N + n0 + n) * E + f;
        const float d = ds[e] + a.dout[((long long)i * a.N + n0 + n) * 2 * E + dir * E + f];
        const float z = a.z[dir][o], cc = a.c[dir][o], sp = a.sprev[dir][o];
        v_dac = d * z * (1.0f - cc * cc);
        v_dz = d * (cc - sp) * z * (1.0f - z);
        v_keep = d * (1.0f - z);
        a.dxi[dir][o] = v_dac;
        a.dxg[dir][((long long)i * a.N + n0 + n) * 2 * E + f] = v_dz;
      }
      dac[e] = v_dac;
      dag[n * 2 * E + f] = v_dz;
      dsn[e] = v_keep;
    }
    __syncthreads();
    // d(rs)[n][k] = sum_j dac[n][j] * Ws[k][j].  Register-tiled when the weights sit in shared memory and the work
    // splits as (k, block of 4 rows) per thread: four j per pass, dac read as broadcast 128-bit values -- 8 shared-
    // memory loads per 16 FMAs instead of 32 (the loops were shared-memory-issue bound: 12 us per encoder step)
    const bool tiled = a.w_in_smem && (E & 3) == 0 && (int)blockDim.x * 4 == ENC_ROWS * E;
    auto rs_epilogue = [&](const int n, const int k, const float acc) {
      const int e = n * E + k;
      const long long o = ((long long)i * a.N + n0 + n) * E + k;
      const float r = a.r[dir][o], sp = a.sprev[dir][o];
      const float dr = acc * sp;
      dsn[e] += acc * r;
      const float v = dr * r * (1.0f - r);
      dag[n * 2 * E + E + k] = v;
      a.dxg[dir][((long long)i * a.N + n0 + n) * 2 * E + E + k] = v;
    };
    if (tiled) {
      const int k = tid % E, nb = (tid / E) * 4;
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < E; j += 4) {
        const float w0 = wsT[j * E + k], w1 = wsT[(j + 1) * E + k], w2 = wsT[(j + 2) * E + k], w3 = wsT[(j + 3) * E + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 dv = *reinterpret_cast<const float4*>(dac + (nb + q) * E + j);
          acc4[q] = fmaf(dv.x, w0, acc4[q]); acc4[q] = fmaf(dv.y, w1, acc4[q]);
          acc4[q] = fmaf(dv.z, w2, acc4[q]); acc4[q] = fmaf(dv.w, w3, acc4[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (nb + q < rows) rs_epilogue(nb + q, k, acc4[q]);
    } else {
      for (int e = tid; e < rows * E; e += blockDim.x) {
        const int n = e / E, k = e % E;
        float acc = 0.0f;
        if (a.w_in_smem) {
#pragma unroll 8
          for (int j = 0; j < E; ++j) acc = fmaf(dac[n * E + j], wsT[j * E + k], acc);
        } else {
          const float* wr = a.Ws[dir] + (long long)k * E;
          for (int j = 0; j < E; ++j) acc = fmaf(dac[n * E + j], wr[j], acc);
        }
        rs_epilogue(n, k, acc);
      }
    }
    __syncthreads();
    // ds[n][k] += sum_j dag[n][j] * Wg[k][j]
    if (tiled) {
      const int k = tid % E, nb = (tid / E) * 4;
      float g4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < 2 * E; j += 4) {
        const float w0 = wgT[j * E + k], w1 = wgT[(j + 1) * E + k], w2 = wgT[(j + 2) * E + k], w3 = wgT[(j + 3) * E + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 dv = *reinterpret_cast<const float4*>(dag + (nb + q) * 2 * E + j);
          g4[q] = fmaf(dv.x, w0, g4[q]); g4[q] = fmaf(dv.y, w1, g4[q]);
          g4[q] = fmaf(dv.z, w2, g4[q]); g4[q] = fmaf(dv.w, w3, g4[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (nb + q < rows) ds[(nb + q) * E + k] = dsn[(nb + q) * E + k] + g4[q];
    } else
    for (int e = tid; e < rows * E; e += blockDim.x) {
      const int n = e / E, k = e % E;
      float acc = 0.0f;
      if (a.w_in_smem) {
#pragma unroll 8
        for (int j = 0; j < 2 * E; ++j) acc = fmaf(dag[n * 2 * E + j], wgT[j * E + k], acc);
      } else {
        const float* wr = a.Wg[dir] + (long long)k * 2 * E;
        for (int j = 0; j < 2 * E; ++j) acc = fmaf(dag[n * 2 * E + j], wr[j], acc);
      }
      ds[e] = dsn[e] + acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < rows * E; e += blockDim.x) {
    const int n = e / E, f = e % E;
    a.ds0[dir][(long long)(n0 + n) * E + f] = ds[e];
  }
}

// ctx[b][u][:] = enc[...] * labels_mask[b][u]   (model.py:645-646); handles the time-axis layout.
// enc is [L][N][C2]; axis0 literal: (L,N) = (B,U) ; axis1: (L,N) = (U,B).
__global__ void context_mask_kernel(const float* __restrict__ enc, const float* __restrict__ lmask, int B,
                                    int U, int C, int time_axis, float* __restrict__ ctx, int backward) {
  const long long n = (long long)B * U * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bu = i / C;
    const int u = (int)(bu % U), b = (int)(bu / U);
    const long long eo = (time_axis == 0) ? i : (((long long)u * B + b) * C + c);
    if (!backward) ctx[i] = enc[eo] * lmask[bu];
    else const_cast<float*>(enc)[eo] = ctx[i] * lmask[bu];  // denc <- dctx * mask
  }
}

// gather rows of a table: dst[m][:] = table[idx[m]][:]
__global__ void gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ idx, long long M,
                                   int F, float* __restrict__ dst) {
  const long long n = M * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / F;
    const int f = (int)(i % F);
    dst[i] = table[(long long)idx[m] * F + f];
  }
}

// dctx[b][u][c] = sum_t phi[t][b][u] * dw[t][b][c]      (one CTA per (b, 32-u tile, 32-c tile))
__global__ void __launch_bounds__(256) dctx_kernel(const float* __restrict__ phi, const float* __restrict__ dw,
                                                   int T, int B, int U, int C, float* __restrict__ dctx) {
  __shared__ float Ps[32][33], Ds[32][33];
  const int b = blockIdx.z, u0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t0 = 0; t0 < T; t0 += 32) {
    for (int j = ty; j < 32; j += 8) {
      const int t = t0 + j;
      Ps[j][tx] = (t < T && u0 + tx < U) ? phi[((long long)t * B + b) * U + u0 + tx] : 0.0f;
      Ds[j][tx] = (t < T && c0 + tx < C) ? dw[((long long)t * B + b) * C + c0 + tx] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float d = Ds[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(Ps[k][ty + 8 * i], d, acc[i]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int u = u0 + ty + 8 * i, c = c0 + tx;
    if (u < U && c < C) dctx[((long long)b * U + u) * C + c] = acc[i];
  }
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// base[b][:] = bias_sum[:] (+ spk[b][:])
__global__ void base_rows_kernel(const float* __restrict__ bias, const float* __restrict__ spk, int B, int F,
                                 float* __restrict__ base) {
  const long long n = (long long)B * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = bias[i % F];
    if (spk) v += spk[i];
    base[i] = v;
  }
}

}  // namespace pb
