// Persistent, warp-specialised tcgen05 kernel for the dominant contraction of a TCN block:
//   q = PReLU(depthwise dilated k3 conv(GroupNorm reg1(p)))      (model/model.py:136/142)
//   racc = (W3 diag(gamma2)) q                                    (model/model.py:144, reg2 folded)
// One CTA per SM loops over 128-frame tiles. Global-memory latency is taken off the compute warps:
//   warps 4,15   raw loaders: cp.async (16 B per lane-op) of the fp32 p tile (+/-4 halo rows) into a 4-stage
//                             shared-memory ring, completion on mbarriers (cp.async.mbarrier.arrive.noinc)
//   warp 14      W loader   : cp.async.bulk of the pre-swizzled fp16 weight image, 2 stages
//   warps 6-13   transform  : raw ring -> GroupNorm/depthwise/PReLU -> fp16 -> 128B-swizzled A operand stage,
//                             statistics of q
//   warp  5      MMA        : tcgen05.mma 128x256x16 (fp16 x fp16 -> fp32) into a DOUBLE-BUFFERED TMEM accumulator
//   warps 0-3    epilogue   : tcgen05.ld -> smem staging -> coalesced stores of the raw accumulators, row sums,
//                             per-utterance column sums; overlaps the next tile's loads / transform / MMA
#include <cstdlib>
#include "kernels.h"
#include "tc_common.cuh"

namespace septfa {

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kHalo = 4;                       // max dilation
constexpr int kRawRows = kTileM + 2 * kHalo;   // 136
constexpr int kRawBytes = kRawRows * 128;      // one K-chunk: 32 in-channels fp32 = 128 B per row
constexpr int kRawStages = 4;
constexpr int kAStage = kTileM * 128;          // 16 KB: 128 rows x 64 halves
constexpr int kWStage = 256 * 128;             // 32 KB: 256 rows x 64 halves
constexpr int kOpStages = 2;
constexpr int kNCH = 8;                        // K = 512 = 8 chunks of 64
constexpr int kStgPitch = 36;
constexpr int kThreadsP = 512;                 // 16 warps
constexpr int kDconvW = 512 * 16 + 512 * 4;

constexpr int kOffA = 0;
constexpr int kOffW = kOffA + kOpStages * kAStage;
constexpr int kOffRaw = kOffW + kOpStages * kWStage;
constexpr int kOffStg = kOffRaw + kRawStages * kRawBytes;
constexpr int kOffWts = kOffStg + 4 * 32 * kStgPitch * 4;
constexpr int kOffAux = kOffWts + kDconvW;
constexpr int kAuxBytesP = 4096;
constexpr int kSmemP = kOffAux + kAuxBytesP + 1024;

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#define TL(role, idx) do { if (p.dbg != nullptr && blockIdx.x == 1 && lane == 0 && (idx) < 256) p.dbg[(role) * 256 + (idx)] = clock64(); } while (0)

__global__ void __launch_bounds__(kThreadsP, 1) k_dconv_persist(DconvParams p, int ntiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffAux);
  uint64_t* raw_full = bars;             // [4]
  uint64_t* raw_empty = bars + 4;        // [4]
  uint64_t* a_full = bars + 8;           // [2]
  uint64_t* w_full = bars + 10;          // [2]
  uint64_t* op_empty = bars + 12;        // [2]
  uint64_t* acc_full = bars + 14;        // [2]
  uint64_t* acc_empty = bars + 16;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float2* tab_a = reinterpret_cast<float2*>(bars + 20);          // [kMaxSegs] mean/rstd of p per segment
  float* seg_acc = reinterpret_cast<float*>(tab_a + kMaxSegs);   // [2*kMaxSegs] slow-path statistics
  float* slots = seg_acc + 2 * kMaxSegs;                         // [8][4]
  float4* w2f_s = reinterpret_cast<float4*>(smem + kOffWts);
  float* c2f_s = reinterpret_cast<float*>(w2f_s + kH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRawStages; ++s) { mbar_init(raw_full + s, 64); mbar_init(raw_empty + s, 8); }
    for (int s = 0; s < kOpStages; ++s) {
      mbar_init(a_full + s, 8);
      mbar_init(w_full + s, 1);
      mbar_init(op_empty + s, 1);
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, 4);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < kH; i += kThreadsP) {
    const int jj = i >> 6, cc = (i >> 3) & 7, oo = i & 7, d = (jj * 8 + oo) * 8 + cc;  // [chunk][output][lane group]
    w2f_s[d] = __ldg(p.w2f + i);
    c2f_s[d] = __ldg(p.c2f + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmine = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (warp == 4 || warp == 15) {
    // ------------------------------------------------------------ raw loaders: cp.async 16 B per lane-op, completion on
    // the stage's mbarrier via cp.async.mbarrier.arrive.noinc (64 arrivals = 2 warps x 32 lanes per phase)
    const int lid = (warp == 4 ? 0 : 32) + lane;   // 0..63
    for (int k = 0; k < nmine; ++k) {
      const int r0 = ((int)blockIdx.x + k * (int)gridDim.x) * kTileM;
      const int lo = max(r0 - kHalo, 0), hi = min(r0 + kTileM + kHalo, p.M);  // valid global rows [lo, hi)
      for (int j = 0; j < kNCH; ++j) {
        const int g = k * kNCH + j, s = g % kRawStages, u = g / kRawStages;
        if (u > 0) mbar_wait(raw_empty + s, (u - 1) & 1, 100 + j);
        const uint32_t dst = smem_u32(smem + kOffRaw + s * kRawBytes);
        const float* src = p.p_in + j * 32;
        // op index i -> (raw row i / 8, 16-byte piece i % 8); 8 consecutive lanes fetch one 128 B row slice
        for (int i = lid; i < kRawRows * 8; i += 64) {
          const int rr = i >> 3, piece = i & 7;
          const int row = r0 - kHalo + rr;
          if (row >= lo && row < hi)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + rr * 128 + piece * 16),
                         "l"(src + (int64_t)row * kC + piece * 4)
                         : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(raw_full + s)) : "memory");
        if (warp == 4) TL(0, g);
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------------------ weight loader
    if (lane == 0) {
      for (int g = 0; g < nmine * kNCH; ++g) {
        const int s = g % kOpStages, u = g / kOpStages, j = g % kNCH;
        if (u > 0) mbar_wait(op_empty + s, (u - 1) & 1, 200 + j);
        mbar_expect_tx(w_full + s, kWStage);
        bulk_copy_g2s(smem + kOffW + s * kWStage, reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)j * kWStage, kWStage,
                      w_full + s);
        TL(1, g);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc_f16(kTileM, 256);
      for (int k = 0; k < nmine; ++k) {
        const int ab = k & 1, au = k >> 1;
        if (au > 0) mbar_wait(acc_empty + ab, (au - 1) & 1, 300);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 256);
        for (int j = 0; j < kNCH; ++j) {
          const int g = k * kNCH + j, s = g % kOpStages, u = g / kOpStages;
          mbar_wait(w_full + s, u & 1, 310 + j);
          TL(5, g);
          mbar_wait(a_full + s, u & 1, 320 + j);
          TL(2, g);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_u32(smem + kOffA + s * kAStage));
          const uint64_t b_desc = make_sw128_desc(smem_u32(smem + kOffW + s * kWStage));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC, (j | kk) != 0);
          umma_commit(op_empty + s);
        }
        umma_commit(acc_full + ab);
      }
    }
    __syncwarp();
  } else if (warp >= 6 && warp < 14) {
    // ------------------------------------------------------------ transform warps
    const int tw = warp - 6, tt = threadIdx.x - 6 * 32;  // 0..255 within the transform group
    const int c8 = lane & 7, rg = lane >> 3;
    for (int k = 0; k < nmine; ++k) {
      const int r0 = ((int)blockIdx.x + k * (int)gridDim.x) * kTileM;
      const int nrows = min(kTileM, p.M - r0);
      const SegMap smap(r0, p.T);
      const int nseg = (r0 + nrows - 1) / p.T - smap.b_first + 1;
      for (int i = tt; i < nseg; i += 256)
        tab_a[i] = stat_mean_rstd(p.st_p + smap.b_first + i, 1.0 / ((double)kC * p.T), 1e-8f);
      for (int i = tt; i < 2 * nseg; i += 256) seg_acc[i] = 0.f;
      named_bar(2, 256);
      SegStat2 qstat;
      for (int j = 0; j < kNCH; ++j) {
        const int g = k * kNCH + j;
        const int rs = g % kRawStages, ru = g / kRawStages;
        const int as = g % kOpStages, au = g / kOpStages;
        mbar_wait(raw_full + rs, ru & 1, 400 + j);
        if (tw == 0) TL(6, g);
        if (au > 0) mbar_wait(op_empty + as, (au - 1) & 1, 410 + j);
        if (tw == 0) TL(3, g);
        const uint8_t* raw = smem + kOffRaw + rs * kRawBytes;
        uint8_t* a_tile = smem + kOffA + as * kAStage;
        const float4* wf = w2f_s + j * 64 + c8;
        const float* cf = c2f_s + j * 64 + c8;
        const int g0 = j * 32 + c8 * 4;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 32 + tw * 4 + rg;
          float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (rl < nrows) {
            const int row = r0 + rl;
            const int sg = smap.seg(row);
            const int t = smap.frame(row, sg);
            const bool okm = t - p.dil >= 0, okp = t + p.dil < p.T;
            const float2 mr = tab_a[sg];
            const float4 xc = *reinterpret_cast<const float4*>(raw + (rl + kHalo) * 128 + c8 * 16);
            float4 xm = make_float4(0.f, 0.f, 0.f, 0.f), xp = xm;
            if (okm) xm = *reinterpret_cast<const float4*>(raw + (rl + kHalo - p.dil) * 128 + c8 * 16);
            if (okp) xp = *reinterpret_cast<const float4*>(raw + (rl + kHalo + p.dil) * 128 + c8 * 16);
            const float vm[4] = {xm.x, xm.y, xm.z, xm.w}, vc[4] = {xc.x, xc.y, xc.z, xc.w}, vp[4] = {xp.x, xp.y, xp.z, xp.w};
            if (okm && okp) {
              const float nmu = -mr.x;
#pragma unroll
              for (int o = 0; o < 8; ++o) {
                const float4 w = wf[o * 8];
                const float conv = fmaf(w.z, vp[o >> 1], fmaf(w.y, vc[o >> 1], w.x * vm[o >> 1]));
                q[o] = prelu(fmaf(mr.y, fmaf(nmu, w.w, conv), cf[o * 8]), p.slope2);
              }
            } else {
              // frames within `dil` of an utterance edge: taps outside are zero padding of the *normalised* signal
              const float4 ga = __ldg(reinterpret_cast<const float4*>(p.g1 + g0));
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.be1 + g0));
              const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float hm = okm ? ((vm[c] - mr.x) * mr.y) * gam[c] + bet[c] : 0.f;
                const float hc = ((vc[c] - mr.x) * mr.y) * gam[c] + bet[c];
                const float hp = okp ? ((vp[c] - mr.x) * mr.y) * gam[c] + bet[c] : 0.f;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const float4 w = __ldg(p.w2b + 2 * (g0 + c) + e);
                  q[2 * c + e] = prelu(w.w + w.x * hm + w.y * hc + w.z * hp, p.slope2);
                }
              }
            }
            qstat.add(sg, q, seg_acc);
          }
          const uint4 pk = make_uint4(pack_half2(q[0], q[1]), pack_half2(q[2], q[3]), pack_half2(q[4], q[5]),
                                      pack_half2(q[6], q[7]));
          *reinterpret_cast<uint4*>(a_tile + sw128_offset(rl, c8)) = pk;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a_full + as);
          mbar_arrive(raw_empty + rs);
        }
        if (tw == 0) TL(4, g);
      }
      // statistics of q for this tile: fixed-order reduction, then double atomics
      qstat.flush_warp(slots, tw);
      named_bar(2, 256);
      for (int i = tt; i < nseg; i += 256) {
        float s = seg_acc[2 * i], qq = seg_acc[2 * i + 1];
        if (i < 2)
          for (int w = 0; w < 8; ++w) { s += slots[w * 4 + 2 * i]; qq += slots[w * 4 + 2 * i + 1]; }
        atomicAdd(&p.st_q[smap.b_first + i].s, (double)s);
        atomicAdd(&p.st_q[smap.b_first + i].ss, (double)qq);
      }
      named_bar(2, 256);  // slots / seg_acc / tab_a are rewritten by the next tile
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp)
    float* stg = reinterpret_cast<float*>(smem + kOffStg) + warp * (32 * kStgPitch);
    for (int k = 0; k < nmine; ++k) {
      const int r0 = ((int)blockIdx.x + k * (int)gridDim.x) * kTileM;
      const int nrows = min(kTileM, p.M - r0);
      const SegMap smap(r0, p.T);
      const int nseg = (r0 + nrows - 1) / p.T - smap.b_first + 1;
      const int ab = k & 1, au = k >> 1;
      mbar_wait(acc_full + ab, au & 1, 500);
      tc_fence_after();
      if (warp == 0) TL(7, 2 * k);
      const int my_rl = warp * 32 + lane;
      float rowacc = 0.f;
      // this lane copies out rows i = it*4 + lane/8 (it = 0..7) of the warp's 32 rows, 4 columns each:
      // per-tile bit masks replace the per-row segment lookups inside the column loop
      const int c4 = (lane & 7) * 4, rsub = lane >> 3;
      uint32_t m_valid = 0, m_seg1 = 0, m_slow = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rl = warp * 32 + it * 4 + rsub;
        if (rl < nrows) {
          const int sg = smap.seg(r0 + rl);
          m_valid |= 1u << it;
          if (sg == 1) m_seg1 |= 1u << it;
          if (sg >= 2) m_slow |= 1u << it;
        }
      }
      const bool any_slow = __any_sync(0xffffffffu, m_slow != 0);  // only when T < 128
      float* orow = p.racc + (int64_t)(r0 + warp * 32 + rsub) * kC + c4;
      for (int cc = 0; cc < 8; ++cc) {
        const int col0 = cc * 32;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ab * 256 + col0), v);
        if (cc == 7) {
          // all TMEM reads of this accumulator buffer are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + ab);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rowacc += v[i];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(stg + lane * kStgPitch + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        __syncwarp();
        float4 ca = make_float4(0.f, 0.f, 0.f, 0.f), c1 = ca;   // column sums: all rows / rows of the 2nd utterance
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if ((m_valid >> it) & 1u) {
            const float4 o = *reinterpret_cast<const float4*>(stg + (it * 4 + rsub) * kStgPitch + c4);
            *reinterpret_cast<float4*>(orow + (int64_t)it * 4 * kC + col0) = o;
            if (!((m_slow >> it) & 1u)) { ca.x += o.x; ca.y += o.y; ca.z += o.z; ca.w += o.w; }
            if ((m_seg1 >> it) & 1u) { c1.x += o.x; c1.y += o.y; c1.z += o.z; c1.w += o.w; }
            if (any_slow && ((m_slow >> it) & 1u)) {
              const int sg = smap.seg(r0 + warp * 32 + it * 4 + rsub);
              double* dst = p.colsum + (size_t)(smap.b_first + sg) * kC + col0 + c4;
              atomicAdd(dst, (double)o.x); atomicAdd(dst + 1, (double)o.y);
              atomicAdd(dst + 2, (double)o.z); atomicAdd(dst + 3, (double)o.w);
            }
          }
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          ca.x += __shfl_xor_sync(0xffffffffu, ca.x, o); ca.y += __shfl_xor_sync(0xffffffffu, ca.y, o);
          ca.z += __shfl_xor_sync(0xffffffffu, ca.z, o); ca.w += __shfl_xor_sync(0xffffffffu, ca.w, o);
        }
        if (nseg > 1) {
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            c1.x += __shfl_xor_sync(0xffffffffu, c1.x, o); c1.y += __shfl_xor_sync(0xffffffffu, c1.y, o);
            c1.z += __shfl_xor_sync(0xffffffffu, c1.z, o); c1.w += __shfl_xor_sync(0xffffffffu, c1.w, o);
          }
        }
        if (lane < 8) {
          double* dst = p.colsum + (size_t)smap.b_first * kC + col0 + c4;
          atomicAdd(dst, (double)(ca.x - c1.x)); atomicAdd(dst + 1, (double)(ca.y - c1.y));
          atomicAdd(dst + 2, (double)(ca.z - c1.z)); atomicAdd(dst + 3, (double)(ca.w - c1.w));
          if (nseg > 1) {
            dst += kC;
            atomicAdd(dst, (double)c1.x); atomicAdd(dst + 1, (double)c1.y);
            atomicAdd(dst + 2, (double)c1.z); atomicAdd(dst + 3, (double)c1.w);
          }
        }
      }
      if (my_rl < nrows) p.rowsum[r0 + my_rl] = rowacc;
      if (warp == 0) TL(7, 2 * k + 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

int g_num_sms = 0;

}  // namespace

cudaError_t dconv_persist_setup() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_dconv_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemP);
}

void launch_dconv_persist(const DconvParams& p, cudaStream_t st) {
  const int ntiles = (p.M + kTileM - 1) / kTileM;
  const int grid = ntiles < g_num_sms ? ntiles : g_num_sms;
  k_dconv_persist<<<grid, kThreadsP, kSmemP, st>>>(p, ntiles);
  ++g_launch_count;
}

}  // namespace septfa
