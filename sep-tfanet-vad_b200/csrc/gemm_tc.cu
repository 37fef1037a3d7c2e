// tcgen05 engine for the three dense contractions of the TCN (the 1x1 convolutions):
//   MODE 0  conv1d      256 -> 256 : A = GN(stream) computed on load;  epilogue +bias, PReLU, stats
//   MODE 1  dconv+conv3 512 -> 256 : A = PReLU(depthwise-dilated-conv(GN1(p))) computed on load
//                                    (q never leaves the SM); epilogue raw accumulators + row/col sums
//   MODE 2  output conv 256 -> 514 : A = GN(PReLU(GN(stream))) on load; epilogue +bias
// One CTA = one 128-frame tile (UMMA M = 128, cta_group::1) x one N tile (256, or 192 x 3 for MODE 2).
// fp16 operands in shared memory (K-major, 128-byte swizzle), fp32 accumulators in TMEM.
//   warps 0-7 : produce the A operand chunk by chunk (global fp32 -> transform -> fp16 -> swizzled
//               st.shared), then run the epilogue (tcgen05.ld -> shared staging -> coalesced stores)
//   warp 8    : streams the pre-swizzled weight image with cp.async.bulk (TMA bulk copy) + mbarrier
//   warp 9    : allocates TMEM, issues tcgen05.mma (one elected thread), tcgen05.commit -> mbarriers
// Reference semantics: model/model.py:130-149 (DepthConv1d), :322-325,357 (TCN.output).
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "kernels.h"
#include "tc_common.cuh"

namespace septfa {

namespace {

constexpr int kTileM = 128;
constexpr int kAChunkBytes = kTileM * 128;  // one K-chunk (64 halves = 128 B) of the A tile
constexpr int kStages = 2;
constexpr int kThreads = 320;
constexpr int kAuxBytes = 4096;
constexpr int kDconvWBytes = 512 * 16 + 512 * 4;  // MODE 1: folded depthwise taps staged in shared memory
constexpr int kStgPitch = 36;               // floats per staged row (32 + 4 pad, 16 B aligned)

struct TcParams {
  int M, T, B;
  const __half* w_img;
  const __half* w_img_lo;   // MODE 2: low part of the fp16 split of the weights
  const void* in;           // [M,256] fp32 (MODE 0/2, and MODE 1 when !H16) or fp16 (MODE 1 with H16)
  StreamNorm norm;
  // MODE 2 prologue
  float slope_o; const Stat2* st_o; const float* g_o; const float* b_o;
  // MODE 1 prologue
  const Stat2* st_p; const float* g1; const float* be1; const float4* w2b; const float4* w2f; const float* c2f;
  float slope2; int dil; Stat2* st_q;
  // epilogue
  const float* bias; float slope;
  void* out; int out_stride;  // fp32, or fp16 for MODE 0/1 with H16
  Stat2* st_out;
  float* rowsum; double* colsum;
  long long* dbg;   // optional timeline buffer (bring-up only)
};

using namespace tc;

// Bring-up timeline (clock64 stamps of one CTA), compiled in only with -DSEPTFA_TIMELINE.
#ifdef SEPTFA_TIMELINE
#define TLG(idx) do { if (p.dbg != nullptr && blockIdx.x == 3 && blockIdx.y == 0 && threadIdx.x == 0 && (idx) < 64) p.dbg[idx] = clock64(); } while (0)
#else
#define TLG(idx) do { } while (0)
#endif

__device__ __forceinline__ float4 ld_half4(const __half* p) {   // 4 consecutive halves (8 B) -> float4
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// H16: the activation tensor exchanged with the neighbouring contraction (conv1's output p = dconv's input; dconv's
// output racc) is stored as fp16 instead of fp32.
template <int MODE, bool H16>
__global__ void __launch_bounds__(kThreads, 2) k_tc_gemm(TcParams p) {
  constexpr int NT = (MODE == 2) ? 192 : 256;
  constexpr int KDIM = (MODE == 1) ? 512 : 256;
  constexpr int NCH = KDIM / 64;
  constexpr int WCH = NT * 128;
  // MODE 2 runs a 3-pass fp16 split (A = A_hi + A_lo, W = W_hi + W_lo; A_hi W_hi + A_lo W_hi + A_hi W_lo): the output
  // conv feeds the VAD head directly and dominated its error budget, and costs < 3 % of the FLOPs.
  constexpr int NSPLIT = (MODE == 2) ? 2 : 1;
  constexpr int STAGE = NSPLIT * (kAChunkBytes + WCH);   // [A_hi][A_lo][W_hi][W_lo]
  constexpr int OFF_W = NSPLIT * kAChunkBytes;
  constexpr uint32_t IDESC = make_idesc_f16(kTileM, NT);

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle; plain pointer arithmetic keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * STAGE);
  uint64_t* full_a = bars;          // [kStages] producers -> MMA
  uint64_t* full_w = bars + 2;      // [kStages] bulk copy -> MMA
  uint64_t* empty = bars + 4;       // [kStages] MMA -> producers / loader
  uint64_t* acc_full = bars + 6;    // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  float2* tab_a = reinterpret_cast<float2*>(bars + 8);
  float2* tab_b = tab_a + kMaxSegs;
  float* seg_acc = reinterpret_cast<float*>(tab_b + kMaxSegs);
  float* rs_x = seg_acc + 2 * kMaxSegs;  // [128] row-sum exchange between the two column halves
  float* slots = rs_x + kTileM;          // [8 warps][4] per-warp statistics partials
  float4* w2f_s = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + kAuxBytes);  // MODE 1: [512] folded taps
  float* c2f_s = reinterpret_cast<float*>(w2f_s + kH);                                       // MODE 1: [512]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * kTileM;
  const int nrows = min(kTileM, p.M - r0);
  const SegMap smap(r0, p.T);
  const int b_first = smap.b_first;
  const int nseg = (r0 + nrows - 1) / p.T - b_first + 1;
  const __half* w_img = p.w_img + (size_t)blockIdx.y * NCH * (WCH / 2);
  const __half* w_img_lo = (NSPLIT == 2) ? p.w_img_lo + (size_t)blockIdx.y * NCH * (WCH / 2) : nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_a + s, 256);
      mbar_init(full_w + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  pdl_launch_dependents();
  if (MODE == 1) {
    // folded depthwise taps (static weights): staged before waiting for the previous kernel
    // staged as [chunk j][output o][lane chunk c8] so that the 8 lane groups of a warp read consecutive words
    for (int i = threadIdx.x; i < kH; i += kThreads) {
      const int jj = i >> 6, cc = (i >> 3) & 7, oo = i & 7, d = (jj * 8 + oo) * 8 + cc;
      w2f_s[d] = __ldg(p.w2f + i);
      c2f_s[d] = __ldg(p.c2f + i);
    }
  }
  pdl_wait();   // everything below reads what earlier kernels of the chain wrote
  {
    const double inv_n = 1.0 / ((double)kC * p.T);
    for (int i = threadIdx.x; i < nseg; i += kThreads) {
      if (MODE == 1) {
        tab_a[i] = stat_mean_rstd(p.st_p + b_first + i, inv_n, 1e-8f);
      } else {
        tab_a[i] = p.norm.gamma != nullptr ? stat_mean_rstd(p.norm.st + b_first + i, p.norm.inv_n, p.norm.eps)
                                           : make_float2(0.f, 1.f);
        if (MODE == 2) tab_b[i] = stat_mean_rstd(p.st_o + b_first + i, inv_n, 1e-5f);
      }
    }
    for (int i = threadIdx.x; i < 2 * nseg; i += kThreads) seg_acc[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  TLG(0);

  if (warp == 8) {
    // ------------------------------------------------------------ weight loader (TMA bulk copies)
    if (lane == 0) {
      for (int j = 0; j < NCH; ++j) {
        const int s = j % kStages, u = j / kStages;
        if (u > 0) mbar_wait(empty + s, (u - 1) & 1, 100 + j);
        mbar_expect_tx(full_w + s, NSPLIT * WCH);
        bulk_copy_g2s(smem + s * STAGE + OFF_W, reinterpret_cast<const uint8_t*>(w_img) + (size_t)j * WCH, WCH, full_w + s);
        if (NSPLIT == 2)
          bulk_copy_g2s(smem + s * STAGE + OFF_W + WCH, reinterpret_cast<const uint8_t*>(w_img_lo) + (size_t)j * WCH, WCH,
                        full_w + s);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      for (int j = 0; j < NCH; ++j) {
        const int s = j % kStages, u = j / kStages;
        mbar_wait(full_w + s, u & 1, 200 + j);
        mbar_wait(full_a + s, u & 1, 300 + j);
        tc_fence_after();
        const uint64_t a_desc = make_sw128_desc(smem_u32(smem + s * STAGE));
        const uint64_t b_desc = make_sw128_desc(smem_u32(smem + s * STAGE + OFF_W));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // UMMA_K = 16 halves = 32 B -> +2 in the (addr >> 4) field
          umma_f16(tmem_base, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC, (j | kk) != 0);
          if (NSPLIT == 2) {
            const uint64_t a_lo = make_sw128_desc(smem_u32(smem + s * STAGE + kAChunkBytes));
            const uint64_t b_lo = make_sw128_desc(smem_u32(smem + s * STAGE + OFF_W + WCH));
            umma_f16(tmem_base, a_lo + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), IDESC, 1u);
            umma_f16(tmem_base, a_desc + (uint64_t)(kk * 2), b_lo + (uint64_t)(kk * 2), IDESC, 1u);
          }
        }
        umma_commit(empty + s);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ A-operand producers (warps 0-7)
    // lane -> (row group rg = lane/8, 16-byte chunk c8 = lane%8); each warp covers 4 rows per step.
    const int c8 = lane & 7, rg = lane >> 3;
    SegStat2 qstat;
    if (MODE == 0) {
      // conv1: A = (x - mean) * rstd (gamma / beta live in the weight image and the bias). The global loads of
      // chunk j+1 are issued before chunk j is converted, so one round trip to HBM is exposed per tile, not per chunk.
      float4 xa[2][4], xb[2][4];
      auto issue = [&](int jj, int buf) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 32 + warp * 4 + rg;
          if (rl < nrows) {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.in) + (int64_t)(r0 + rl) * kC + jj * 64 + c8 * 8);
            xa[buf][it] = __ldg(src);
            xb[buf][it] = __ldg(src + 1);
          } else {
            xa[buf][it] = xb[buf][it] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      float2 mrs[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rl = it * 32 + warp * 4 + rg;
        mrs[it] = rl < nrows ? tab_a[smap.seg(r0 + rl)] : make_float2(0.f, 0.f);
      }
      issue(0, 0);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int s = j % kStages, u = j / kStages;
        if (j + 1 < NCH) issue(j + 1, (j + 1) & 1);
        if (u > 0) mbar_wait(empty + s, (u - 1) & 1, 400 + j);
        uint8_t* a_tile = smem + s * STAGE;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 32 + warp * 4 + rg;
          const float4 x0 = xa[j & 1][it], x1 = xb[j & 1][it];
          const float sc = mrs[it].y, nb = -mrs[it].x * mrs[it].y;
          const uint4 pk = make_uint4(pack_half2(fmaf(x0.x, sc, nb), fmaf(x0.y, sc, nb)), pack_half2(fmaf(x0.z, sc, nb), fmaf(x0.w, sc, nb)),
                                      pack_half2(fmaf(x1.x, sc, nb), fmaf(x1.y, sc, nb)), pack_half2(fmaf(x1.z, sc, nb), fmaf(x1.w, sc, nb)));
          *reinterpret_cast<uint4*>(a_tile + sw128_offset(rl, c8)) = pk;
        }
        fence_proxy_async();
        mbar_arrive(full_a + s);
        TLG(1 + j);
      }
    } else if (MODE == 1) {
      // dconv + res_out: q channels 64j + 8*c8 .. +7  <-  in-channels g0 .. g0+3 (out-channel o reads in-channel o/2).
      // GroupNorm reg1 is folded into the taps (w2f = w * gamma, c2f = b2 + beta * sum(w)):
      //   q = PReLU(c2f + rstd * (sum_k w2f[k] p[t+(k-1)d] - mean * sum_k w2f[k]))      (all taps inside the utterance)
      // A K-chunk is produced in two steps of 2 row groups; the global loads of step st+1 are issued before step st is
      // computed (register double buffer), so the round trip to L2/HBM overlaps the arithmetic instead of being paid
      // twice per chunk (measured: 5700 -> cycles per chunk).
      using RawT = typename std::conditional<H16, uint2, float4>::type;
      struct Step { RawT xm[2], xc[2], xp[2]; int flg[2]; };   // flg: segment | (t-d) ok << 8 | (t+d) ok << 9 | row valid << 10
      auto ldraw = [](const void* base, int64_t off) -> RawT {
        if constexpr (H16) return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + off));
        else return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
      };
      auto to4 = [](const RawT& v, float (&o)[4]) {
        if constexpr (H16) {
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
          o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
        } else {
          o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
      };
      auto issue = [&](int st, Step& sp) {
        const int jj = st >> 1, half = st & 1, g0 = jj * 32 + c8 * 4;
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          const int rl = (half * 2 + i2) * 32 + warp * 4 + rg;
          sp.flg[i2] = 0;
          if (rl < nrows) {
            const int row = r0 + rl;
            const int sg = smap.seg(row);
            const int t = smap.frame(row, sg);
            int f = sg | 1024;
            const int64_t off = (int64_t)row * kC + g0;
            sp.xc[i2] = ldraw(p.in, off);
            if (t - p.dil >= 0) { f |= 256; sp.xm[i2] = ldraw(p.in, off - (int64_t)p.dil * kC); }
            if (t + p.dil < p.T) { f |= 512; sp.xp[i2] = ldraw(p.in, off + (int64_t)p.dil * kC); }
            sp.flg[i2] = f;
          }
        }
      };
      auto compute = [&](int st, const Step& sp) {
        const int jj = st >> 1, half = st & 1, g0 = jj * 32 + c8 * 4;
        uint8_t* a_tile = smem + (jj % kStages) * STAGE;
        const float4* wf = w2f_s + jj * 64 + c8;   // folded taps of this lane's 8 output channels: wf[o * 8]
        const float* cf = c2f_s + jj * 64 + c8;
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          const int rl = (half * 2 + i2) * 32 + warp * 4 + rg;
          float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (sp.flg[i2] & 1024) {
            const int sg = sp.flg[i2] & 255;
            const bool okm = sp.flg[i2] & 256, okp = sp.flg[i2] & 512;
            const float2 mr = tab_a[sg];
            float vm[4] = {0.f, 0.f, 0.f, 0.f}, vc[4], vp[4] = {0.f, 0.f, 0.f, 0.f};
            to4(sp.xc[i2], vc);
            if (okm) to4(sp.xm[i2], vm);
            if (okp) to4(sp.xp[i2], vp);
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
              for (int g = 0; g < 4; ++g) {
                const float hm = okm ? ((vm[g] - mr.x) * mr.y) * gam[g] + bet[g] : 0.f;
                const float hc = ((vc[g] - mr.x) * mr.y) * gam[g] + bet[g];
                const float hp = okp ? ((vp[g] - mr.x) * mr.y) * gam[g] + bet[g] : 0.f;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const float4 w = __ldg(p.w2b + 2 * (g0 + g) + e);
                  q[2 * g + e] = prelu(w.w + w.x * hm + w.y * hc + w.z * hp, p.slope2);
                }
              }
            }
            qstat.add(sg, q, seg_acc);
          }
          const uint4 pk = make_uint4(pack_half2(q[0], q[1]), pack_half2(q[2], q[3]), pack_half2(q[4], q[5]),
                                      pack_half2(q[6], q[7]));
          *reinterpret_cast<uint4*>(a_tile + sw128_offset(rl, c8)) = pk;
        }
      };
      Step sa, sb;
      issue(0, sa);
#pragma unroll 1
      for (int j = 0; j < NCH; ++j) {
        const int s = j % kStages, u = j / kStages;
        issue(2 * j + 1, sb);
        if (u > 0) mbar_wait(empty + s, (u - 1) & 1, 400 + j);
        compute(2 * j, sa);
        if (j + 1 < NCH) issue(2 * j + 2, sa);
        compute(2 * j + 1, sb);
        fence_proxy_async();
        mbar_arrive(full_a + s);
        TLG(1 + j);
      }
    } else
    for (int j = 0; j < NCH; ++j) {
      const int s = j % kStages, u = j / kStages;
      if (u > 0) mbar_wait(empty + s, (u - 1) & 1, 400 + j);
      uint8_t* a_tile = smem + s * STAGE;
      {  // MODE 2 (MODE 0 and 1 have their own software-pipelined loops above)
        const int kc = j * 64 + c8 * 8;
        const bool has_norm = p.norm.gamma != nullptr;
        float ga[8], be[8], go[8], bo[8];
        if (MODE == 2) {
          if (has_norm) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { ga[i] = __ldg(p.norm.gamma + kc + i); be[i] = __ldg(p.norm.beta + kc + i); }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { go[i] = __ldg(p.g_o + kc + i); bo[i] = __ldg(p.b_o + kc + i); }
        }
        float4 x0[4], x1[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 32 + warp * 4 + rg;
          if (rl < nrows) {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.in) + (int64_t)(r0 + rl) * kC + kc);
            x0[it] = __ldg(src);
            x1[it] = __ldg(src + 1);
          } else {
            x0[it] = x1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 32 + warp * 4 + rg;
          float y[8] = {x0[it].x, x0[it].y, x0[it].z, x0[it].w, x1[it].x, x1[it].y, x1[it].z, x1[it].w};
          if (rl < nrows) {
            const int sg = smap.seg(r0 + rl);
            const float2 mr = tab_a[sg];
            if (MODE == 0) {
              // gamma / beta of the stream norm live in the weight image and the bias: A = (x - mean) * rstd
              const float nb = -mr.x * mr.y;
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = fmaf(y[i], mr.y, nb);
            } else {
              if (has_norm) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = ((y[i] - mr.x) * mr.y) * ga[i] + be[i];
              }
              const float2 mo = tab_b[sg];
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = ((prelu(y[i], p.slope_o) - mo.x) * mo.y) * go[i] + bo[i];
            }
          }
          const uint4 pk = make_uint4(pack_half2(y[0], y[1]), pack_half2(y[2], y[3]), pack_half2(y[4], y[5]),
                                      pack_half2(y[6], y[7]));
          *reinterpret_cast<uint4*>(a_tile + sw128_offset(rl, c8)) = pk;
          if (NSPLIT == 2) {
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = y[i] - __half2float(__float2half_rn(y[i]));   // exact residual in fp32
            const uint4 pl = make_uint4(pack_half2(r[0], r[1]), pack_half2(r[2], r[3]), pack_half2(r[4], r[5]),
                                        pack_half2(r[6], r[7]));
            *reinterpret_cast<uint4*>(a_tile + kAChunkBytes + sw128_offset(rl, c8)) = pl;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(full_a + s);
      TLG(1 + j);
    }
    if (MODE == 1) qstat.flush_warp(slots, warp);

    // ------------------------------------------------------------ epilogue (warps 0-7)
    // warp w reads TMEM lanes 32*(w%4).. (rows) and columns (w/4)*NT/2 .. in chunks of 32.
    TLG(10);
    mbar_wait(acc_full, 0, 500);
    tc_fence_after();
    TLG(11);
    const int lq = warp & 3, ch = warp >> 2;
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * kStgPitch);  // aliases the (now idle) stage buffers
    const int my_rl = lq * 32 + lane;       // the row this thread owns in TMEM
    float rowacc = 0.f;
    SegStat2 ostat;
    constexpr int NCC = NT / 64;            // 32-column chunks per column half
    for (int cc = 0; cc < NCC; ++cc) {
      const int col0 = ch * (NT / 2) + cc * 32;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)col0, v);
      TLG(12 + cc * 4);
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) rowacc += v[i];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stg + lane * kStgPitch + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      __syncwarp();
      TLG(13 + cc * 4);
      // coalesced copy-out: 8 lanes x float4 = one 128 B row segment, 4 rows per instruction
      const int c4 = (lane & 7) * 4;
      const int gcol = (MODE == 2 ? (int)blockIdx.y * NT : 0) + col0 + c4;
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE != 1) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
      float4 cs0 = make_float4(0.f, 0.f, 0.f, 0.f), cs1 = cs0;   // column sums of the tile's 1st / 2nd utterance
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int i = it * 4 + (lane >> 3);
        const int rl = lq * 32 + i;
        if (rl < nrows) {
          float4 o = *reinterpret_cast<const float4*>(stg + i * kStgPitch + c4);
          const int row = r0 + rl;
          const int sg = smap.seg(row);
          if (MODE == 0) {
            o.x = prelu(o.x + bias4.x, p.slope); o.y = prelu(o.y + bias4.y, p.slope);
            o.z = prelu(o.z + bias4.z, p.slope); o.w = prelu(o.w + bias4.w, p.slope);
            const float ov[4] = {o.x, o.y, o.z, o.w};
            ostat.add(sg, ov, seg_acc);
          } else if (MODE == 2) {
            o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
          } else {
            if (sg == 0) { cs0.x += o.x; cs0.y += o.y; cs0.z += o.z; cs0.w += o.w; }
            else if (sg == 1) { cs1.x += o.x; cs1.y += o.y; cs1.z += o.z; cs1.w += o.w; }
            else {  // only when T < 128
              double* dst = p.colsum + (size_t)(b_first + sg) * kC + col0 + c4;
              atomicAdd(dst, (double)o.x); atomicAdd(dst + 1, (double)o.y);
              atomicAdd(dst + 2, (double)o.z); atomicAdd(dst + 3, (double)o.w);
            }
          }
          if (H16 && MODE != 2)
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + (int64_t)row * p.out_stride + gcol) =
                make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
          else
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (int64_t)row * p.out_stride + gcol) = o;
        }
      }
      TLG(14 + cc * 4);
      if (MODE == 1) {
        // column sums: fold the 4 row groups of the warp, then one global atomic per column and utterance
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          cs0.x += __shfl_xor_sync(0xffffffffu, cs0.x, o); cs0.y += __shfl_xor_sync(0xffffffffu, cs0.y, o);
          cs0.z += __shfl_xor_sync(0xffffffffu, cs0.z, o); cs0.w += __shfl_xor_sync(0xffffffffu, cs0.w, o);
          cs1.x += __shfl_xor_sync(0xffffffffu, cs1.x, o); cs1.y += __shfl_xor_sync(0xffffffffu, cs1.y, o);
          cs1.z += __shfl_xor_sync(0xffffffffu, cs1.z, o); cs1.w += __shfl_xor_sync(0xffffffffu, cs1.w, o);
        }
        if (lane < 8) {
          double* dst = p.colsum + (size_t)b_first * kC + col0 + c4;
          atomicAdd(dst, (double)cs0.x); atomicAdd(dst + 1, (double)cs0.y);
          atomicAdd(dst + 2, (double)cs0.z); atomicAdd(dst + 3, (double)cs0.w);
          if (nseg > 1) {
            dst += kC;
            atomicAdd(dst, (double)cs1.x); atomicAdd(dst + 1, (double)cs1.y);
            atomicAdd(dst + 2, (double)cs1.z); atomicAdd(dst + 3, (double)cs1.w);
          }
        }
      }
    }
    TLG(30);
    if (MODE == 0) ostat.flush_warp(slots, warp);
    if (MODE == 1) {
      // row sums: the two column halves (warps w and w+4) own the same rows
      if (ch == 1) rs_x[my_rl] = rowacc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0 && my_rl < nrows) p.rowsum[r0 + my_rl] = rowacc + rs_x[my_rl];
    }
  }

  tc_fence_before();
  __syncthreads();
  TLG(31);
  if (warp == 9) tmem_dealloc(tmem_base, 256);
  Stat2* sdst = (MODE == 0) ? p.st_out : (MODE == 1 ? p.st_q : nullptr);
  if (sdst != nullptr && blockIdx.y == 0) seg_stats_commit(slots, 8, seg_acc, nseg, sdst + b_first);
}

template <int MODE, bool H16>
void launch_mode(const TcParams& p, int ntiles_n, cudaStream_t st) {
  constexpr int NT = (MODE == 2) ? 192 : 256;
  constexpr int smem = kStages * (MODE == 2 ? 2 : 1) * (kAChunkBytes + NT * 128) + kAuxBytes + 1024 + (MODE == 1 ? kDconvWBytes : 0);
  dim3 grid((p.M + kTileM - 1) / kTileM, ntiles_n);
  launch_k(k_tc_gemm<MODE, H16>, grid, dim3(kThreads), smem, st, true, p);
}

}  // namespace

long long* g_tl_conv1 = nullptr;  // bring-up timeline of one conv1 launch (SEPTFA_TIMELINE)

template <int MODE, bool H16>
cudaError_t setup_one(int smem) {
  cudaFuncSetAttribute(k_tc_gemm<MODE, H16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return cudaFuncSetAttribute(k_tc_gemm<MODE, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

cudaError_t tc_gemm_setup() {
  // two CTAs per SM need (almost) the whole shared-memory carveout
  const int s0 = kStages * (kAChunkBytes + 256 * 128) + kAuxBytes + 1024;
  const int s2 = kStages * 2 * (kAChunkBytes + 192 * 128) + kAuxBytes + 1024;
  cudaError_t e;
  if ((e = setup_one<0, false>(s0)) != cudaSuccess) return e;
  if ((e = setup_one<0, true>(s0)) != cudaSuccess) return e;
  if ((e = setup_one<1, false>(s0 + kDconvWBytes)) != cudaSuccess) return e;
  if ((e = setup_one<1, true>(s0 + kDconvWBytes)) != cudaSuccess) return e;
  return setup_one<2, false>(s2);
}

void launch_tc_conv1(const Conv1Params& c, cudaStream_t st) {
  TcParams p{};
  p.M = c.M; p.T = c.T; p.B = c.B;
  p.w_img = c.w_img; p.in = c.w_in; p.norm = c.norm;
  p.bias = c.bias_f; p.slope = c.slope;
  p.out = c.p_out; p.out_stride = kC; p.st_out = c.st_p;
  p.dbg = g_tl_conv1;
  if (c.half_io) launch_mode<0, true>(p, 1, st); else launch_mode<0, false>(p, 1, st);
}

void launch_tc_dconv(const DconvParams& c, cudaStream_t st) {
  TcParams p{};
  p.M = c.M; p.T = c.T; p.B = c.B;
  p.w_img = c.w_img; p.in = c.p_in;
  p.st_p = c.st_p; p.g1 = c.g1; p.be1 = c.be1; p.w2b = c.w2b; p.w2f = c.w2f; p.c2f = c.c2f;
  p.slope2 = c.slope2; p.dil = c.dil; p.st_q = c.st_q;
  p.out = c.racc; p.out_stride = kC; p.rowsum = c.rowsum; p.colsum = c.colsum;
  p.dbg = c.dbg;
  if (c.half_io) launch_mode<1, true>(p, 1, st); else launch_mode<1, false>(p, 1, st);
}

void launch_tc_outconv(const OutConvParams& c, cudaStream_t st) {
  TcParams p{};
  p.M = c.M; p.T = c.T; p.B = c.B;
  p.w_img = c.w_img; p.w_img_lo = c.w_img_lo; p.in = c.w_in; p.norm = c.norm;
  p.slope_o = c.slope_o; p.st_o = c.st_o; p.g_o = c.g_o; p.b_o = c.b_o;
  p.bias = c.bias;
  p.out = c.logits; p.out_stride = kLogitStride;
  launch_mode<2, false>(p, 3, st);
}

}  // namespace septfa
