// Persistent, warp-specialised tcgen05 kernel for the second half of a TCN block (model/model.py:136/142,144):
//   q = PReLU(depthwise dilated k3 conv(GN1(p))),  racc = (W3 diag(g2)) q,  statistics of q, row / column sums of racc
// with BOTH contractions on the tensor cores. The depthwise convolution is block-diagonal (output channel o reads input
// channel o/2), so for a group of 16 input channels and one tap k it is a tiny GEMM
//   D1[128 frames x 32 outputs] += P_k[128 x 16] * B_k[16 x 32],   P_k = the p tile shifted by (k-1)*dilation frames,
// whose A operand needs no data movement at all: p is stored by conv1 as per-K-group planes [32][rows][8 channels]
// (16 bytes per frame and plane), one TMA bulk copy brings a plane's 136 frames (tile + halo) into shared memory already
// in the canonical no-swizzle K-major core-matrix layout (8 frames x 16 B), and a shift by s frames is the same buffer
// addressed 16*s bytes further on. GroupNorm reg1 is folded into the taps (w' = w * gamma, fp16) and into a per-row affine
//   x = rstd * conv + (c2f - rstd * mean * sw),   q = PReLU(x)
// applied to the fp32 accumulator on its way from tensor memory to the fp16 A operand of the res_out GEMM - 9 SASS
// instructions per PAIR of q values (FFMA2 pipe) instead of ~34 for the CUDA-core depthwise producer of gemm_tc.cu
// (MODE 1), which bounded that kernel at 20 % of the tensor roofline.
// Zero padding of the NORMALISED signal (the reference pads h1, not p): the tensor-core convolution reads whatever
// neighbours the shifted view holds; the <= 2*dil rows per utterance whose tap falls outside are corrected exactly,
//   x_true = x_mma - w'_k * (rstd * (p_wrong - mean) + beta / gamma),
// by the lane that owns the row (the same fp16 weight and fp16 input the tensor core multiplied).
//
// One CTA per SM walks 128-frame tiles; roles (16 warps):
//   warp 0        p loader   : 4 bulk copies (K-group planes) per 32-channel chunk, ring of 3
//   warp 1        W loader   : pre-swizzled res_out weight image chunks (32 KB), ring of 2. Optionally (option
//                              "dconv_cluster" = 2) the CTAs run as clusters of two in lock step and rank 0 fetches every
//                              chunk once with a MULTICAST bulk copy into both CTAs' shared memory, a stage being refilled
//                              when both CTAs' MMAs have released it (multicast commit). Measured: no gain (1.28 vs 1.25 ms
//                              per 24 launches) - what bounds the kernel is the L2 -> SM fabric (3.3 KB/clk = 6.4 TB/s over
//                              148 SMs, 80 % of it this weight stream), and a 2-CTA multicast still delivers every byte to
//                              every SM; halving the bytes per SM needs the 2-SM MMA (cta_group::2), see DESIGN.md
//   warp 2        MMA issuer : per chunk 6 mini-MMAs (2 groups x 3 taps, M128 N32 K16, both operands in shared memory) into
//                              the ring, issued kLookahead chunks ahead of the 4 main MMAs (M128 N256 K16, A from TMEM) into D2
//   warps 4-11    transform  : tcgen05.ld D1 -> affine, PReLU, statistics, fp16 -> tcgen05.st back into the SAME tensor-memory
//                              buffer: q never touches shared memory, the res_out MMA takes its A operand from TMEM
//   warps 12-15   epilogue   : tcgen05.ld D2 (two loads in flight) -> fp16 rows staged per lane -> D2 released -> per-utterance
//                              column sums from the staged tile, row sums, one TMA bulk store per row
// Tensor memory: D2 = columns 0-255; ring of 4 x 64 columns at 256-511: a buffer first holds D1 (fp32, 64 columns), then the
// packed q (fp16 pairs: columns +0..15 and +32..47, each transform warp writing inside the half it has just read), and is
// handed back by the res_out MMA's commit. The tap matrices (48 KB) stay resident in shared memory.
// Synchronisation is mbarriers only. Requires T >= 128 (a tile touches at most two utterances) and the plane layout of
// p (conv1's persistent kernel writes it); otherwise launch_tc_dconv's kernel is used.
#include <algorithm>
#include <cuda.h>
#include "kernels.h"
#include "tc_common.cuh"

namespace septfa {

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kHalo = kPlaneHalo;                   // 4 = the largest dilation
constexpr int kSlabRows = kTileM + 2 * kHalo;       // 136 frames: tile + halo
constexpr int kPlaneBytes = kSlabRows * 16;         // 2176: one K-group (8 channels) of the slab
constexpr int kPChunkBytes = 4 * kPlaneBytes;       // 8704: 32 input channels
constexpr int kPStages = 4;
constexpr int kWBytes = 256 * 128;                  // 32 KB: one K-chunk (64) of the res_out weight image
constexpr int kWStages = 2;
constexpr int kD1Bufs = 4;
constexpr int kLookahead = 2;                       // chunks the depthwise mini-GEMMs run ahead of the res_out GEMM
constexpr int kEpiPitch = 512 + 16;                 // bytes per staged row (256 fp16 columns + pad: conflict-free STS.128)
constexpr int kEpiWarpBytes = 32 * kEpiPitch;
constexpr int kOffTap = 0;
constexpr int kOffW = kOffTap + kDconvTapBytes;               // 1024-aligned (128B swizzle)
constexpr int kOffP = kOffW + kWStages * kWBytes;
constexpr int kOffEpi = kOffP + kPStages * kPChunkBytes;
constexpr int kOffSwc = kOffEpi + 4 * kEpiWarpBytes;
constexpr int kOffBar = kOffSwc + 4096;
constexpr int kSmemBytes = kOffBar + 512;
constexpr int kThreadsD = 16 * 32;
static_assert(kOffW % 1024 == 0 && kSmemBytes <= 232448, "shared-memory plan");

struct DmParams {
  alignas(64) CUtensorMap w_tmap;   // res_out weight image as a 2-D tensor [2048 rows][64 halves], box = one K-chunk
  int use_tmap;
  int M, T, ntiles, Mp, dil;
  float slope2;
  const __half* p_planes;      // [32 K-groups][Mp slots][8 channels]; frame r lives in slot r + kHalo
  const Stat2* st_p;
  const uint8_t* tap_img;      // [16 groups][3 taps][32 x 16 fp16, no-swizzle K-major]
  const float4* swc;           // [256]: {sw[2i], sw[2i+1], c2f[2i], c2f[2i+1]}
  const float* w16;            // [3][512] the fp16-rounded folded taps as fp32 (edge corrections)
  const float* bog;            // [256] beta1 / gamma1
  const __half* w_img;         // res_out: 8 K-chunks x [256 rows x 128 B], 128B-swizzled
  __half* racc;                // [M,256]
  float* rowsum; double* colsum; Stat2* st_q;
  int desc_swap;               // bring-up switch: exchange the leading / stride byte offsets of the no-swizzle descriptors
  int cluster;                 // CTAs per cluster (1, or 2: multicast weight stream)
};

// Shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between core matrices adjacent in K, SBO = between core matrices adjacent in M / N.
__device__ __forceinline__ uint64_t make_ns_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Bring-up timeline (clock64 stamps of CTA 0's roles), compiled in only with -DSEPTFA_DM_TIMELINE.
#ifdef SEPTFA_DM_TIMELINE
__device__ long long g_dm_tl[10][64];
#define DTL(role, idx) do { if (blockIdx.x == 0 && (idx) < 64) g_dm_tl[role][idx] = clock64(); } while (0)
#else
#define DTL(role, idx) do { } while (0)
#endif

// TMA tensor load of a 2-D box (SASS: UTMALDG), completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

template <bool AMAX>
__global__ void __launch_bounds__(kThreadsD, 1) k_dconv_mma(const __grid_constant__ DmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* w_full = bars + 6;        // [2]
  uint64_t* w_empty = bars + 8;       // [2] MMA commit
  uint64_t* d1_full = bars + 10;      // [4] MMA commit
  uint64_t* d1_empty = bars + 14;     // [4] res_out MMA commit (the buffer held D1, then q)
  uint64_t* a2_full = bars + 18;      // [4] 256 transform threads (q is in tensor memory)
  uint64_t* d2_full = bars + 22;      // MMA commit
  uint64_t* d2_empty = bars + 23;     // 128 epilogue threads
  uint64_t* tap_full = bars + 24;
  uint64_t* p_full = bars + 25;       // [4] bulk-copy bytes
  uint64_t* p_empty = bars + 29;      // [4] MMA commit + one lane of each transform warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  constexpr uint32_t IDESC_MAIN = make_idesc_f16(kTileM, 256);
  constexpr uint32_t IDESC_MINI = make_idesc_f16(kTileM, 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    for (int s = 0; s < kPStages; ++s) { mbar_init(p_full + s, 1); mbar_init(p_empty + s, 9); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, p.cluster); }   // every CTA of the cluster releases a stage
    for (int s = 0; s < kD1Bufs; ++s) { mbar_init(d1_full + s, 1); mbar_init(d1_empty + s, 1); mbar_init(a2_full + s, 256); }
    mbar_init(d2_full, 1); mbar_init(d2_empty, 128); mbar_init(tap_full, 1);
    fence_mbar_init();
    // static weights: the tap matrices go to shared memory once, before the dependency wait
    mbar_expect_tx(tap_full, kDconvTapBytes);
    for (int i = 0; i < 3; ++i) bulk_copy_g2s(smem + kOffTap + i * 16384, p.tap_img + i * 16384, 16384, tap_full);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x < 256) reinterpret_cast<float4*>(smem + kOffSwc)[threadIdx.x] = __ldg(p.swc + threadIdx.x);
  pdl_wait();   // everything below reads what earlier kernels of the chain wrote
  tc_fence_before();
  cluster_sync_all();   // (also a CTA barrier) the partner's mbarriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DTL(9, 0);
  // the pair walks tile pairs: cluster c takes pairs c, c + nclusters, ...; rank r the tile 2 * pair + r. Both CTAs run
  // the same number of tiles (the last pair of an odd tile count has an empty tile: no valid rows, nothing stored).
  const int cs = p.cluster;                      // 1 or 2
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const int npairs = (p.ntiles + cs - 1) / cs;
  const int first_pair = (int)blockIdx.x / cs, pair_stride = (int)gridDim.x / cs;
  const int my_tiles = first_pair < npairs ? (npairs - first_pair + pair_stride - 1) / pair_stride : 0;
  const int first = cs * first_pair + (int)crank, stride = cs * pair_stride, tile_end = cs * npairs;

  if (warp == 0) {
    // ------------------------------------------------------------ p loader (K-group planes, tile + halo)
    if (lane == 0) {
      int g = 0;
      for (int tile = first; tile < tile_end; tile += stride) {
        for (int j = 0; j < 8; ++j, ++g) {
          const int s = g % kPStages, u = g / kPStages;
          if (u > 0) mbar_wait(p_empty + s, (u - 1) & 1, 100 + j);
          DTL(0, g);
          mbar_expect_tx(p_full + s, kPChunkBytes);
#pragma unroll
          for (int kg = 0; kg < 4; ++kg)   // frames r0 - 4 .. r0 + 131 = slots r0 .. r0 + 135
            bulk_copy_g2s(smem + kOffP + s * kPChunkBytes + kg * kPlaneBytes,
                          p.p_planes + ((size_t)(j * 4 + kg) * p.Mp + (size_t)tile * kTileM) * 8, kPlaneBytes, p_full + s);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ res_out weight loader
    if (lane == 0) {
      int g = 0;
      for (int tile = first; tile < tile_end; tile += stride) {
        for (int j = 0; j < 8; ++j, ++g) {
          const int s = g % kWStages, u = g / kWStages;
          if (u > 0) mbar_wait(w_empty + s, (u - 1) & 1, 150 + j);   // both CTAs' MMAs are done with the stage
          DTL(1, g);
          mbar_expect_tx(w_full + s, kWBytes);
          if (cs == 1 && p.use_tmap)   // tensor-map TMA load of the chunk's 256 x 64 box (measured: same rate as the linear copy)
            tma_load_2d(smem + kOffW + s * kWBytes, &p.w_tmap, 0, j * 256, w_full + s);
          else if (cs == 1)
            bulk_copy_g2s(smem + kOffW + s * kWBytes, reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)j * kWBytes, kWBytes, w_full + s);
          else if ((uint32_t)(j & 1) == crank)   // the two CTAs' TMA engines take turns: each issues half of the stream
            bulk_copy_g2s_mc(smem + kOffW + s * kWBytes, reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)j * kWBytes, kWBytes, w_full + s,
                             (uint16_t)3);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(tap_full, 0, 190);
      const int total = my_tiles * 8;
      const uint32_t lbo_a = p.desc_swap ? 128u : (uint32_t)kPlaneBytes, sbo_a = p.desc_swap ? (uint32_t)kPlaneBytes : 128u;
      const uint32_t lbo_b = p.desc_swap ? 128u : 512u, sbo_b = p.desc_swap ? 512u : 128u;
      for (int g = 0; g < total + kLookahead; ++g) {
        if (g < total) {
          // depthwise mini-GEMMs of chunk g -> D1[g % 4]
          const int sp = g % kPStages, up = g / kPStages, b = g % kD1Bufs, ub = g / kD1Bufs, j = g & 7;
          mbar_wait(p_full + sp, up & 1, 200 + j);
          if (ub > 0) mbar_wait(d1_empty + b, (ub - 1) & 1, 210 + j);
          tc_fence_after();
          DTL(2, g);
          const uint32_t slab = smem_u32(smem + kOffP + sp * kPChunkBytes);
#pragma unroll
          for (int grp = 0; grp < 2; ++grp) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const uint32_t a_addr = slab + (uint32_t)(grp * 2 * kPlaneBytes + (kHalo + (k - 1) * p.dil) * 16);
              const uint32_t b_addr = smem_u32(smem + kOffTap + ((j * 2 + grp) * 3 + k) * 1024);
              umma_f16(tmem_base + 256u + (uint32_t)(b * 64 + grp * 32), make_ns_desc(a_addr, lbo_a, sbo_a),
                       make_ns_desc(b_addr, lbo_b, sbo_b), IDESC_MINI, k != 0);
            }
          }
          umma_commit(d1_full + b);
          umma_commit(p_empty + sp);
        }
        if (g >= kLookahead) {
          // res_out GEMM of chunk gm into D2
          const int gm = g - kLookahead, j = gm & 7, lt = gm >> 3;
          if (j == 0 && lt > 0) { mbar_wait(d2_empty, (lt - 1) & 1, 220); tc_fence_after(); }
          const int ba = gm % kD1Bufs, ua = gm / kD1Bufs, sw = gm % kWStages, uw = gm / kWStages;
          mbar_wait(w_full + sw, uw & 1, 230 + j);
          DTL(8, gm);
          mbar_wait(a2_full + ba, ua & 1, 240 + j);
          tc_fence_after();
          DTL(3, gm);
          const uint32_t a_tmem = tmem_base + 256u + (uint32_t)(ba * 64);   // q: K 0..31 at +0..15, K 32..63 at +32..47
          const uint64_t b_desc = make_sw128_desc(smem_u32(smem + kOffW + sw * kWBytes));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ts(tmem_base, a_tmem + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), b_desc + (uint64_t)(kk * 2), IDESC_MAIN, (j | kk) != 0);
          umma_commit(d1_empty + ba);
          if (cs == 1) umma_commit(w_empty + sw); else umma_commit_mc(w_empty + sw, (uint16_t)3);
          if (j == 7) umma_commit(d2_full);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------ transform: D1 -> q -> A2
    const int q4 = warp & 3, hf = (warp - 4) >> 2;       // TMEM lane quarter (hardware: warp % 4), column half of the chunk
    const int rl = q4 * 32 + lane;                        // a lane = a row of the tile
    const float2 sl2 = make_float2(p.slope2, p.slope2);
    const float4* swc_s = reinterpret_cast<const float4*>(smem + kOffSwc);
    const double inv_n = 1.0 / ((double)kC * p.T);
    int g = 0;
    for (int tile = first; tile < tile_end; tile += stride) {
      const int r0 = tile * kTileM, nrows = max(0, min(kTileM, p.M - r0));
      const int b_first = r0 / p.T, e1 = (b_first + 1) * p.T;   // first row of the tile's second utterance
      float2 mr = make_float2(0.f, 1.f);
      if (nrows > 0 && lane < 2 && (lane == 0 || e1 < r0 + nrows)) mr = stat_mean_rstd(p.st_p + b_first + lane, inv_n, 1e-8f);
      const float m0 = __shfl_sync(0xffffffffu, mr.x, 0), s0 = __shfl_sync(0xffffffffu, mr.y, 0);
      const float m1 = __shfl_sync(0xffffffffu, mr.x, 1), s1 = __shfl_sync(0xffffffffu, mr.y, 1);
      const int row = r0 + rl;
      const bool valid = rl < nrows, second = row >= e1;
      const float mean = second ? m1 : m0, rstd = second ? s1 : s0;
      const int t = row - (second ? e1 : b_first * p.T);
      const bool lo_inv = valid && t - p.dil < 0, hi_inv = valid && t + p.dil >= p.T;
      const float2 rs2 = make_float2(rstd, rstd), nr2 = make_float2(-mean * rstd, -mean * rstd);
      float2 accS = make_float2(0.f, 0.f), accQ = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int j = 0; j < 8; ++j, ++g) {
        const int b = g % kD1Bufs, ub = g / kD1Bufs, sp = g % kPStages;
        mbar_wait(d1_full + b, ub & 1, 300 + j);
        tc_fence_after();
        if (warp == 4 && lane == 0) DTL(4, g);
        float v[32];
        const uint32_t t_buf = tmem_base + 256u + (uint32_t)(b * 64 + hf * 32) + ((uint32_t)(q4 * 32) << 16);
        tmem_ld32(t_buf, v);
        if (warp == 4 && lane == 0) DTL(7, g);
        if (lo_inv || hi_inv) {
          // zero padding of the normalised signal: take back what the out-of-utterance tap contributed
          mbar_wait(p_full + sp, (g / kPStages) & 1, 310 + j);   // (completed long ago: orders our reads after the bulk copy)
          const float inv_rstd = 1.0f / rstd;
          const uint8_t* slab = smem + kOffP + sp * kPChunkBytes;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (side == 0 ? lo_inv : hi_inv) {
              const int srow = kHalo + rl + (side == 0 ? -p.dil : p.dil);
              const float* wk = p.w16 + (side == 0 ? 0 : 2 * kH) + j * 64 + hf * 32;
#pragma unroll
              for (int kg2 = 0; kg2 < 2; ++kg2) {
                const uint4 raw = *reinterpret_cast<const uint4*>(slab + (hf * 2 + kg2) * kPlaneBytes + srow * 16);
                const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const float2 pv = __half22float2(*reinterpret_cast<const __half2*>(&rw[e2]));
                  const int c = kg2 * 8 + e2 * 2;     // input channel within this lane's 16
                  const float2 bg = __ldg(reinterpret_cast<const float2*>(p.bog + j * 32 + hf * 16 + c));
                  const float u0 = (pv.x - mean) + bg.x * inv_rstd, u1 = (pv.y - mean) + bg.y * inv_rstd;
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(wk + 2 * c));
                  v[2 * c] = fmaf(-w4.x, u0, v[2 * c]);
                  v[2 * c + 1] = fmaf(-w4.y, u0, v[2 * c + 1]);
                  v[2 * c + 2] = fmaf(-w4.z, u1, v[2 * c + 2]);
                  v[2 * c + 3] = fmaf(-w4.w, u1, v[2 * c + 3]);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(p_empty + sp);
        const float4* sw = swc_s + j * 32 + hf * 16;
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 t4 = sw[i * 4 + k];    // same address in every lane: broadcast
            const float2 k0 = __ffma2_rn(nr2, make_float2(t4.x, t4.y), make_float2(t4.z, t4.w));
            const float2 x = __ffma2_rn(rs2, make_float2(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]), k0);
            const float2 ax = __fmul2_rn(sl2, x);
            float2 qv;
            if constexpr (AMAX) qv = make_float2(fmaxf(x.x, ax.x), fmaxf(x.y, ax.y));
            else qv = make_float2(fminf(x.x, ax.x), fminf(x.y, ax.y));
            accS = __fadd2_rn(accS, qv);
            accQ = __ffma2_rn(qv, qv, accQ);
            h[i * 4 + k] = pack_half2(qv.x, qv.y);
          }
        }
        if (warp == 4 && lane == 0 && g < 32) DTL(7, 32 + g);
        tmem_st16(t_buf, h);           // q (fp16 pairs) over the first 16 of the 32 columns this warp has just read
        tc_fence_before();
        mbar_arrive(a2_full + b);
        if (warp == 4 && lane == 0) DTL(5, g);
      }
      // statistics of q of this warp's 32 rows x 256 columns, per utterance: fixed-order shuffle trees, double atomics
      const float sv = valid ? accS.x + accS.y : 0.f, qv = valid ? accQ.x + accQ.y : 0.f;
      const float a0 = warp_sum(second ? 0.f : sv), c0 = warp_sum(second ? 0.f : qv);
      const float a1 = warp_sum(second ? sv : 0.f), c1 = warp_sum(second ? qv : 0.f);
      if (lane == 0 && nrows > 0) {
        atomicAdd(&p.st_q[b_first].s, (double)a0);
        atomicAdd(&p.st_q[b_first].ss, (double)c0);
        if (e1 < r0 + nrows) {
          atomicAdd(&p.st_q[b_first + 1].s, (double)a1);
          atomicAdd(&p.st_q[b_first + 1].ss, (double)c1);
        }
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------ epilogue: D2 -> racc (fp16), row / column sums
    const int q4 = warp & 3;
    const int rl = q4 * 32 + lane;
    uint8_t* stg_w = smem + kOffEpi + (warp - 12) * kEpiWarpBytes;
    uint8_t* stg = stg_w + lane * kEpiPitch;
    int lt = 0;
    for (int tile = first; tile < tile_end; tile += stride, ++lt) {
      const int r0 = tile * kTileM, nrows = max(0, min(kTileM, p.M - r0));
      const int b_first = r0 / p.T, e1 = (b_first + 1) * p.T;
      const int row = r0 + rl;
      const bool valid = rl < nrows;
      const int nv_w = max(0, min(32, nrows - q4 * 32));               // valid rows of this warp
      const int n0_w = max(0, min(nv_w, e1 - (r0 + q4 * 32)));         // ... that belong to the tile's first utterance
      mbar_wait(d2_full, lt & 1, 500);
      tc_fence_after();
      if (warp == 12 && lane == 0) DTL(6, lt * 4);
      bulk_wait_read_all();            // this lane's previous row has left the staging buffer
      __syncwarp();
      // drain D2 with two tensor-memory loads in flight; fp32 -> fp16 into this lane's staged row
      float2 racc2 = make_float2(0.f, 0.f);
      uint32_t va[32], vb[32];
      const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16);
      auto pack_store = [&](const uint32_t (&u)[32], int cc) {
        uint4* dst = reinterpret_cast<uint4*>(stg + cc * 64);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v8[e] = __uint_as_float(u[8 * i + e]);
          racc2 = __fadd2_rn(racc2, __fadd2_rn(__fadd2_rn(make_float2(v8[0], v8[1]), make_float2(v8[2], v8[3])),
                                               __fadd2_rn(make_float2(v8[4], v8[5]), make_float2(v8[6], v8[7]))));
          dst[i] = make_uint4(pack_half2(v8[0], v8[1]), pack_half2(v8[2], v8[3]), pack_half2(v8[4], v8[5]), pack_half2(v8[6], v8[7]));
        }
      };
      tmem_ld32_nowait(t_row, va);
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        tmem_ld_wait();                                              // va (chunk 2 c2) has landed
        tmem_ld32_nowait(t_row + (uint32_t)((2 * c2 + 1) * 32), vb);
        pack_store(va, 2 * c2);
        tmem_ld_wait();                                              // vb (chunk 2 c2 + 1)
        if (c2 < 3) tmem_ld32_nowait(t_row + (uint32_t)((2 * c2 + 2) * 32), va);
        else {                         // the whole accumulator is in registers / shared memory: hand D2 back to the MMA warp
          tc_fence_before();
          mbar_arrive(d2_empty);
          if (warp == 12 && lane == 0) DTL(6, lt * 4 + 1);
        }
        pack_store(vb, 2 * c2 + 1);
      }
      if (valid) p.rowsum[row] = racc2.x + racc2.y;
      __syncwarp();
      if (warp == 12 && lane == 0) DTL(6, 32 + lt * 4);
      // per-utterance column sums of this warp's rows (from the stored fp16 values): lane -> 8 columns
      float2 c0[4], c1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) c0[e] = c1[e] = make_float2(0.f, 0.f);
      for (int r = 0; r < nv_w; ++r) {
        const uint4 raw = *reinterpret_cast<const uint4*>(stg_w + r * kEpiPitch + lane * 16);
        const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
        if (r < n0_w) {
#pragma unroll
          for (int e = 0; e < 4; ++e) c0[e] = __fadd2_rn(c0[e], __half22float2(*reinterpret_cast<const __half2*>(&rw[e])));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) c1[e] = __fadd2_rn(c1[e], __half22float2(*reinterpret_cast<const __half2*>(&rw[e])));
        }
      }
      if (warp == 12 && lane == 0) DTL(6, 32 + lt * 4 + 1);
      double* cdst = p.colsum + (size_t)b_first * kC + lane * 8;
      if (n0_w > 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { atomicAdd(cdst + 2 * e, (double)c0[e].x); atomicAdd(cdst + 2 * e + 1, (double)c0[e].y); }
      }
      if (nv_w > n0_w) {
        cdst += kC;
#pragma unroll
        for (int e = 0; e < 4; ++e) { atomicAdd(cdst + 2 * e, (double)c1[e].x); atomicAdd(cdst + 2 * e + 1, (double)c1[e].y); }
      }
      if (warp == 12 && lane == 0) DTL(6, 32 + lt * 4 + 2);
      fence_proxy_async();             // our generic-proxy stores -> visible to the bulk copy
      __syncwarp();                    // every lane's column-sum reads of the staging buffer are done
      if (valid) {
        bulk_copy_s2g(p.racc + (size_t)row * kC, stg, 512);
        bulk_commit_group();
      }
      if (warp == 12 && lane == 0) DTL(6, lt * 4 + 2);
    }
    bulk_wait_read_all();
  }

  tc_fence_before();
  cluster_sync_all();   // neither CTA leaves while the other may still multicast into it
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int g_dm_sm_count = 0;

}  // namespace

#ifdef SEPTFA_DM_TIMELINE
extern "C" int septfa_debug_dm_timeline(long long* out) {   // [10][64] clock64 stamps of CTA 0 of the last launch
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_dm_tl, sizeof(long long) * 640) == cudaSuccess ? 0 : -1;
}
#endif

cudaError_t dconv_mma_setup() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_dm_sm_count, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(k_dconv_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_dconv_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

void launch_dconv_mma(const DconvMmaParams& c, cudaStream_t st) {
  DmParams p{};
  p.M = c.M; p.T = c.T; p.ntiles = (c.M + kTileM - 1) / kTileM; p.Mp = c.Mp; p.dil = c.dil; p.slope2 = c.slope2;
  p.p_planes = c.p_planes; p.st_p = c.st_p; p.tap_img = c.tap_img; p.swc = c.swc; p.w16 = c.w16; p.bog = c.bog;
  p.w_img = c.w_img; p.racc = c.racc; p.rowsum = c.rowsum; p.colsum = c.colsum; p.st_q = c.st_q;
  p.desc_swap = ctx().dconv_desc_swap;
  p.use_tmap = (c.w_tmap != nullptr && ctx().dconv_w_tmap) ? 1 : 0;
  if (p.use_tmap) p.w_tmap = *reinterpret_cast<const CUtensorMap*>(c.w_tmap);
  p.cluster = ctx().dconv_cluster == 2 ? 2 : 1;
  const int npairs = (p.ntiles + p.cluster - 1) / p.cluster;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.cluster * std::min(g_dm_sm_count / p.cluster, npairs));
  cfg.blockDim = dim3(kThreadsD);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ctx().use_pdl ? 2 : 1;
  if (c.slope2 <= 1.f) cudaLaunchKernelEx(&cfg, k_dconv_mma<true>, p);
  else cudaLaunchKernelEx(&cfg, k_dconv_mma<false>, p);
  ++ctx().launches;
}

}  // namespace septfa
