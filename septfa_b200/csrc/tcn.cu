// TCN element-wise / reduction kernels (TF-attention gates, post-block GroupNorm residual
// updates) and the plain-fp32 CUDA-core engine for the 1x1 convolutions.
// Reference: model/model.py:130-149 (DepthConv1d), :197-208 (TF_Attention), :343-357 (TCN.forward).
#include "kernels.h"

namespace septfa {

// ------------------------------------------------------------------------------------------
// Segment table: the rows of a tile are consecutive frames of >= 1 utterances ("segments").
// seg_tab[i] = {mean, rstd} of the tile's i-th utterance for one statistics buffer.
__device__ __forceinline__ void fill_seg_table(float2* tab, const Stat2* st, double inv_n, float eps, int b_first,
                                               int nseg) {
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) tab[i] = stat_mean_rstd(st + b_first + i, inv_n, eps);
}

constexpr int kRowsPerCta = 128;  // rows per CTA of the element-wise kernels

struct RowCtx {
  int r0, nrows, b_first, nseg;
};
__device__ __forceinline__ RowCtx row_ctx(int M, int T) {
  RowCtx c;
  c.r0 = blockIdx.x * kRowsPerCta;
  c.nrows = min(kRowsPerCta, M - c.r0);
  c.b_first = c.r0 / T;
  c.nseg = (c.r0 + c.nrows - 1) / T - c.b_first + 1;
  return c;
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8_plain(const float* p, float (&v)[8]) {  // coherent path (buffer is rewritten in place)
  const float4 a = reinterpret_cast<const float4*>(p)[0];
  const float4 b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// MODE: 0 = statistics of v only; 1 = apply (w <- y + GN(v)), plus statistics of the new stream
// when the ln mode is recursive.  Per element (coefficients cached per utterance in registers):
//   y = w * Ay + By            Ay = rstd_y * gamma_y, By = beta_y - mean_y * Ay      (stream norm)
//   r*g = gt[row] * (racc * G1 + G2)      G1 = ra * gf, G2 = rb * gf             (TF-attention gates)
//   v = y + r*g (recursive) | r*g (residual);  out = y + v * Av + Bv   (Av, Bv: ln_first / ln_modules)
template <int MODE, bool H16>
__global__ void __launch_bounds__(256, 4) k_resid(ResidParams p) {
  __shared__ float2 tab_y[kMaxSegs];   // stream norm
  __shared__ float2 tab_v[kMaxSegs];   // stats of v (MODE 1)
  __shared__ float acc_sm[kMaxSegs * 2];
  __shared__ float slots[8 * 4];
  pdl_launch_dependents();
  pdl_wait();
  const RowCtx c = row_ctx(p.M, p.T);
  const bool has_norm = p.norm.gamma != nullptr;
  const bool use_v_norm = MODE == 1 && p.mode != LN_NONE;
  if (has_norm) fill_seg_table(tab_y, p.norm.st, p.norm.inv_n, p.norm.eps, c.b_first, c.nseg);
  if (use_v_norm) fill_seg_table(tab_v, p.st_v, 1.0 / (256.0 * p.T), 1e-5f, c.b_first, c.nseg);
  for (int i = threadIdx.x; i < c.nseg * 2; i += blockDim.x) acc_sm[i] = 0.f;
  __syncthreads();

  // a warp covers half a row (lane = 4 channels); warps (2k, 2k+1) walk the same rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = ((warp & 1) * 32 + lane) * 4;
  float4 Ay, By, G1, G2, Av, Bv;
  Av = Bv = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = -1;
  SegStat2 acc;
  const SegMap smap(c.r0, p.T);
  auto ld4 = [](const float* q) { return __ldg(reinterpret_cast<const float4*>(q)); };
  auto load_coeffs = [&](int sg) {
    const int b = c.b_first + sg;
    if (has_norm) {
      const float2 my = tab_y[sg];
      const float4 g = ld4(p.norm.gamma + c0), be = ld4(p.norm.beta + c0);
      Ay = make_float4(my.y * g.x, my.y * g.y, my.y * g.z, my.y * g.w);
      By = make_float4(be.x - my.x * Ay.x, be.y - my.x * Ay.y, be.z - my.x * Ay.z, be.w - my.x * Ay.w);
    } else {
      Ay = make_float4(1.f, 1.f, 1.f, 1.f);
      By = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float ra = __ldg(p.ra + b);
    const float4 rb = ld4(p.rb + b * kC + c0), gf = ld4(p.gf + b * kC + c0);
    G1 = make_float4(ra * gf.x, ra * gf.y, ra * gf.z, ra * gf.w);
    G2 = make_float4(rb.x * gf.x, rb.y * gf.y, rb.z * gf.z, rb.w * gf.w);
    if (use_v_norm) {
      const float2 mv = tab_v[sg];
      const float4 g = ld4(p.g_a + c0), be = ld4(p.b_a + c0);
      Av = make_float4(mv.y * g.x, mv.y * g.y, mv.y * g.z, mv.y * g.w);
      Bv = make_float4(be.x - mv.x * Av.x, be.y - mv.x * Av.y, be.z - mv.x * Av.z, be.w - mv.x * Av.w);
    }
    cur = sg;
  };
  constexpr int RS = 2;  // rows per warp step: 4 x 16 B loads in flight per lane, 32 warps per SM
  for (int i = (warp >> 1) * RS; i < c.nrows; i += 4 * RS) {
    float4 w[RS], ra4[RS];
    float gt[RS];
#pragma unroll
    for (int k = 0; k < RS; ++k) {
      if (i + k < c.nrows) {
        const int row = c.r0 + i + k;
        w[k] = *reinterpret_cast<const float4*>(p.w + (int64_t)row * kC + c0);  // coherent: rewritten in place
        if (H16) {
          const uint2 hv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.racc) + (int64_t)row * kC + c0));
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&hv.x));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
          ra4[k] = make_float4(f0.x, f0.y, f1.x, f1.y);
        } else {
          ra4[k] = ld4(reinterpret_cast<const float*>(p.racc) + (int64_t)row * kC + c0);
        }
        gt[k] = __ldg(p.gt + row);
      }
    }
#pragma unroll
    for (int k = 0; k < RS; ++k) {
      if (i + k < c.nrows) {
        const int row = c.r0 + i + k;
        const int sg = smap.seg(row);
        if (sg != cur) load_coeffs(sg);
        const float wv[4] = {w[k].x, w[k].y, w[k].z, w[k].w}, rv[4] = {ra4[k].x, ra4[k].y, ra4[k].z, ra4[k].w};
        const float ay[4] = {Ay.x, Ay.y, Ay.z, Ay.w}, by[4] = {By.x, By.y, By.z, By.w};
        const float g1[4] = {G1.x, G1.y, G1.z, G1.w}, g2[4] = {G2.x, G2.y, G2.z, G2.w};
        float v[4], y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[j] = fmaf(wv[j], ay[j], by[j]);
          const float r = gt[k] * fmaf(rv[j], g1[j], g2[j]);
          v[j] = (p.mode == LN_RECURSIVE) ? y[j] + r : r;
        }
        if (MODE == 0) {
          acc.add(sg, v, acc_sm);
        } else {
          const float av[4] = {Av.x, Av.y, Av.z, Av.w}, bv[4] = {Bv.x, Bv.y, Bv.z, Bv.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = (p.mode == LN_NONE) ? y[j] + v[j] : y[j] + fmaf(v[j], av[j], bv[j]);
          *reinterpret_cast<float4*>(p.w + (int64_t)row * kC + c0) = make_float4(o[0], o[1], o[2], o[3]);
          if (p.mode == LN_RECURSIVE) acc.add(sg, o, acc_sm);
        }
      }
    }
  }
  acc.flush_warp(slots, warp);
  __syncthreads();
  Stat2* dst = (MODE == 0) ? p.st_v : p.st_w;
  if (dst != nullptr && (MODE == 0 || p.mode == LN_RECURSIVE)) seg_stats_commit(slots, 8, acc_sm, c.nseg, dst + c.b_first);
}

void launch_resid_stats(const ResidParams& p, cudaStream_t st) {
  const int grid = (p.M + kRowsPerCta - 1) / kRowsPerCta;
  if (p.racc_half) launch_k(k_resid<0, true>, dim3(grid), dim3(256), 0, st, true, p);
  else launch_k(k_resid<0, false>, dim3(grid), dim3(256), 0, st, true, p);
}
void launch_resid_apply(const ResidParams& p, cudaStream_t st) {
  const int grid = (p.M + kRowsPerCta - 1) / kRowsPerCta;
  if (p.racc_half) launch_k(k_resid<1, true>, dim3(grid), dim3(256), 0, st, true, p);
  else launch_k(k_resid<1, false>, dim3(grid), dim3(256), 0, st, true, p);
}

// Statistics of PReLU(y) for the output GroupNorm (model.py:322-323,357).
__global__ void __launch_bounds__(256) k_out_stats(const float* __restrict__ w, StreamNorm norm, float slope, int M,
                                                   int T, Stat2* __restrict__ st_o) {
  __shared__ float2 tab_y[kMaxSegs];
  __shared__ float acc_sm[kMaxSegs * 2];
  __shared__ float slots[8 * 4];
  pdl_launch_dependents();
  pdl_wait();
  const RowCtx c = row_ctx(M, T);
  const bool has_norm = norm.gamma != nullptr;
  if (has_norm) fill_seg_table(tab_y, norm.st, norm.inv_n, norm.eps, c.b_first, c.nseg);
  for (int i = threadIdx.x; i < c.nseg * 2; i += blockDim.x) acc_sm[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c0 = lane * 8;
  float gy[8], by[8];
  if (has_norm) { ld8(norm.gamma + c0, gy); ld8(norm.beta + c0, by); }
  SegStat2 acc;
  const SegMap smap(c.r0, T);
  for (int i = warp; i < c.nrows; i += 8) {
    const int row = c.r0 + i;
    const int sg = smap.seg(row);
    float wv[8], z[8];
    ld8(w + (int64_t)row * kC + c0, wv);
    const float2 my = has_norm ? tab_y[sg] : make_float2(0.f, 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = prelu(has_norm ? ((wv[j] - my.x) * my.y) * gy[j] + by[j] : wv[j], slope);
    acc.add(sg, z, acc_sm);
  }
  acc.flush_warp(slots, warp);
  __syncthreads();
  seg_stats_commit(slots, 8, acc_sm, c.nseg, st_o + c.b_first);
}

void launch_out_stats(const float* w, StreamNorm norm, float slope, int M, int T, Stat2* st_o, cudaStream_t st) {
  launch_k(k_out_stats, dim3((M + kRowsPerCta - 1) / kRowsPerCta), dim3(256), 0, st, true, w, norm, slope, M, T, st_o);
}

// ------------------------------------------------------------------------------------------
// TF_Attention gates (model.py:197-208) from the row/column sums of the raw conv3 accumulators,
// plus the per-utterance affine that turns raw accumulators into r (GroupNorm reg2 folded).
// One CTA per utterance.
__device__ __forceinline__ float tf_chain(const float* m, int n, int j, const float* w1, float b1, const float* w2,
                                          float b2, float slope) {
  // u1 = conv(k3,p1,d1)(m); u2 = conv(k3,p2,d2)(u1); both zero-pad their own input.
  float u2 = b2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = j + 2 * (k - 1);
    if (i < 0 || i >= n) continue;
    float u1 = b1;
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int ii = i + kk - 1;
      if (ii >= 0 && ii < n) u1 += w1[kk] * m[ii];
    }
    u2 += w2[k] * u1;
  }
  return sigmoidf_acc(prelu(u2, slope));
}

__global__ void __launch_bounds__(256) k_tf_gate(GateParams p) {
  float* sm = p.mt + (int64_t)blockIdx.x * p.T;   // m_t [T] of this utterance: global scratch, so T is unbounded
  __shared__ float m_f[kC];
  __shared__ float red[32];
  __shared__ float s_ra, s_mu, s_rbmean;
  const int b = blockIdx.x, tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (tid == 0) {
    const float2 mr = stat_mean_rstd(p.st_q + b, 1.0 / ((double)kH * p.T), 1e-8f);
    s_mu = mr.x;
    s_ra = mr.y;
    p.ra[b] = mr.y;
  }
  __syncthreads();
  const float ra = s_ra, mu = s_mu;
  // thread = channel
  const float rbv = __ldg(p.c03 + tid) - ra * mu * __ldg(p.s3 + tid);
  p.rb[b * kC + tid] = rbv;
  m_f[tid] = ra * (float)(__ldg(p.colsum + b * kC + tid) / (double)p.T) + rbv;
  float s = warp_sum(rbv);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_rbmean = t / (float)kC;
  }
  __syncthreads();
  const float rbmean = s_rbmean;
  for (int t = tid; t < p.T; t += 256) sm[t] = ra * (__ldg(p.rowsum + (int64_t)b * p.T + t) / (float)kC) + rbmean;
  __syncthreads();
  if (p.tf.enabled) {
    p.gf[b * kC + tid] = tf_chain(m_f, kC, tid, p.tf.wf1, p.tf.bf1, p.tf.wf2, p.tf.bf2, p.tf.af);
    for (int t = tid; t < p.T; t += 256)
      p.gt[(int64_t)b * p.T + t] = tf_chain(sm, p.T, t, p.tf.wt1, p.tf.bt1, p.tf.wt2, p.tf.bt2, p.tf.at);
  } else {
    p.gf[b * kC + tid] = 1.f;
    for (int t = tid; t < p.T; t += 256) p.gt[(int64_t)b * p.T + t] = 1.f;
  }
}

void launch_tf_gate(const GateParams& p, cudaStream_t st) {
  launch_k(k_tf_gate, dim3(p.B), dim3(256), 0, st, true, p);
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core engine: one CTA per frame, thread = output channel. Used for bring-up and as
// the on-device full-precision cross-check of the tcgen05 engine (SEPTFA_ENGINE_FP32_SIMT).
__global__ void __launch_bounds__(256) k_ref_conv1(Conv1Params p) {
  __shared__ float y[kC];
  __shared__ float red[64];
  const int row = blockIdx.x, b = row / p.T, n = threadIdx.x;
  float v = __ldg(p.w_in + (int64_t)row * kC + n);
  if (p.norm.gamma != nullptr) {
    const float2 mr = stat_mean_rstd(p.norm.st + b, p.norm.inv_n, p.norm.eps);
    v = ((v - mr.x) * mr.y) * __ldg(p.norm.gamma + n) + __ldg(p.norm.beta + n);
  }
  y[n] = v;
  __syncthreads();
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < kC; ++k) acc = fmaf(y[k], __ldg(p.w_t + k * kC + n), acc);
  const float o = prelu(acc + __ldg(p.bias + n), p.slope);
  reinterpret_cast<float*>(p.p_out)[(int64_t)row * kC + n] = o;
  block_stat_atomic(o, o * o, p.st_p + b, red);
}

__global__ void __launch_bounds__(256) k_ref_dconv(DconvParams p) {
  __shared__ float q[kH];
  __shared__ float red[64];
  const int row = blockIdx.x, b = row / p.T, t = row - b * p.T, c = threadIdx.x;
  const float2 mr = stat_mean_rstd(p.st_p + b, 1.0 / ((double)kC * p.T), 1e-8f);
  const float g = __ldg(p.g1 + c), be = __ldg(p.be1 + c);
  float h[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int tt = t + (k - 1) * p.dil;
    h[k] = (tt >= 0 && tt < p.T) ? ((__ldg(reinterpret_cast<const float*>(p.p_in) + (int64_t)(row + (k - 1) * p.dil) * kC + c) - mr.x) * mr.y) * g + be
                                 : 0.f;
  }
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float4 w = __ldg(p.w2b + 2 * c + j);
    const float v = prelu(w.w + w.x * h[0] + w.y * h[1] + w.z * h[2], p.slope2);
    q[2 * c + j] = v;
    s += v;
    ss += v * v;
  }
  block_stat_atomic(s, ss, p.st_q + b, red);  // contains __syncthreads
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < kH; ++k) acc = fmaf(q[k], __ldg(p.w_t + k * kC + c), acc);
  reinterpret_cast<float*>(p.racc)[(int64_t)row * kC + c] = acc;
  atomicAdd(p.colsum + b * kC + c, (double)acc);
  float rs = warp_sum(acc);
  if ((c & 31) == 0) red[c >> 5] = rs;
  __syncthreads();
  if (c == 0) {
    float tsum = 0.f;
    for (int i = 0; i < 8; ++i) tsum += red[i];
    p.rowsum[row] = tsum;
  }
}

__global__ void __launch_bounds__(256) k_ref_outconv(OutConvParams p) {
  __shared__ float z[kC];
  const int row = blockIdx.x, b = row / p.T, c = threadIdx.x;
  float v = __ldg(p.w_in + (int64_t)row * kC + c);
  if (p.norm.gamma != nullptr) {
    const float2 mr = stat_mean_rstd(p.norm.st + b, p.norm.inv_n, p.norm.eps);
    v = ((v - mr.x) * mr.y) * __ldg(p.norm.gamma + c) + __ldg(p.norm.beta + c);
  }
  const float2 mo = stat_mean_rstd(p.st_o + b, 1.0 / ((double)kC * p.T), 1e-5f);
  z[c] = ((prelu(v, p.slope_o) - mo.x) * mo.y) * __ldg(p.g_o + c) + __ldg(p.b_o + c);
  __syncthreads();
  for (int n = c; n < kLogitStride; n += 256) {
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < kC; ++k) acc = fmaf(z[k], __ldg(p.w_t + k * kLogitStride + n), acc);
    p.logits[(int64_t)row * kLogitStride + n] = acc + __ldg(p.bias + n);
  }
}

void launch_ref_conv1(const Conv1Params& p, cudaStream_t st) {
  k_ref_conv1<<<p.M, 256, 0, st>>>(p);
  ++ctx().launches;
}
void launch_ref_dconv(const DconvParams& p, cudaStream_t st) {
  k_ref_dconv<<<p.M, 256, 0, st>>>(p);
  ++ctx().launches;
}
void launch_ref_outconv(const OutConvParams& p, cudaStream_t st) {
  k_ref_outconv<<<p.M, 256, 0, st>>>(p);
  ++ctx().launches;
}

}  // namespace septfa
