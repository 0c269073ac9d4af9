// conv_tile.cuh -- spatially tiled contraction: TMA-staged halo tiles + implicit im2col through the
// UMMA shared-memory descriptors.
//
// Same job as conv_umma.cuh (reference genXMatrix_kernel cbconv2d_cg_backend.cu:138-161 + the cuBLAS
// GEMM conv2d_cg.py:342-349 + updateOutput_kernel cbconv2d_cg_backend.cu:175-189), for change sets
// that are spatially clustered (blocks, dense frames): instead of gathering every receptive-field
// tap of every changed pixel (kH*kW copies of each state pixel), the work unit is an 8 x 16 tile of
// output pixels that contains at least one changed pixel (tile list built by cb_dilate_compact_tiles):
//
//   * ONE cp.async.bulk.tensor.4d (TMA, tile mode, out-of-image = zero fill) per operand plane
//     stages the (8+kW-1) x (16+kH-1) halo of the pixel-major state in shared memory -- every state
//     pixel crosses L2 -> SM once per tile, not once per tap;
//   * the A operand of tcgen05.mma is never materialised: an 8-row core group of the canonical
//     K-major layout = 8 x-consecutive halo pixels (row pitch = pixel bytes = swizzle width:
//     16 B none / 32 B / 64 B / 128 B), group stride SBO = halo row pitch, and the descriptor's
//     start address is simply the tap's pixel -- the hardware swizzle is a function of the absolute
//     shared-memory address, so any 16/32/64/128-byte aligned start reads TMA-swizzled data
//     correctly (probed on the B200: tools/umma_tile_probe.cu).  16-byte pixels (<= 8 sixteen-bit
//     channels, the RGB input layer) take two taps per K=16 instruction with LBO = tap distance;
//     pixels wider than 128 B are staged as 128-byte channel blocks;
//   * weights stream through a TMA ring exactly as in conv_umma.cuh (or stay resident when the whole
//     filter bank fits the ring), accumulators are double-buffered in TMEM, the epilogue writes
//     only the rows whose bit is set in the dilated change bitmap (bias / ReLU / convert / one
//     contiguous channel run per pixel) -- untouched pixels keep their exact previous bits.
//
// Warp roles: 0 weight TMA, 1 halo TMA, 2 MMA issuer (+ TMEM alloc), 3 spare, 4.. epilogue.
#pragma once
#include "conv_umma.cuh"

namespace cb {

constexpr int TL_W = 8, TL_H = 16;             // output tile (pixels): 128 rows of the M dimension
constexpr int TL_NHALO = 2;                    // halo buffers (tile i+1 loads while tile i computes; 1 if smem is tight)
constexpr int TL_MAXB = 12;                    // weight ring slots
constexpr int TL_MAXTAB = 1024;                // K-step table entries (8 KB)
constexpr int TL_CTRL_BYTES = 512;

__host__ __device__ inline int tile_grid_y(int H) { return (H + TL_H - 1) / TL_H; }
__host__ __device__ inline int tile_grid_x(int W) { return (W + TL_W - 1) / TL_W; }
// tile row stride: a 32-pixel bitmap word covers 4 tiles, kept inside one group of four
__host__ __device__ inline int tile_grid_xp(int W) { return ((W + 31) / 32) * 4; }
// tile workspace (int32 words): [0] append counter, [1] number of dirty tiles of the last
// cb_dilate_compact_tiles, [2..3] pad, then NT epoch stamps, then the tile list (NT entries)
__host__ __device__ inline size_t tile_ws_words(int B, int H, int W) {
  return 4 + 2 * (size_t)B * tile_grid_y(H) * tile_grid_xp(W);
}

struct TileCtrl {
  uint64_t b_full[TL_MAXB], b_empty[TL_MAXB];
  uint64_t halo_full[TL_NHALO], halo_empty[TL_NHALO];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base, pad;
};
static_assert(sizeof(TileCtrl) <= TL_CTRL_BYTES, "ctrl block too large");

struct TileGeom {
  int B, H, W, Wd, TY, TXp;      // images, bitmap row words, tile grid
  int kH, kW, HWX, HWY;          // filter, halo extent (pixels)
  int Cp, Kp, num_kb;            // operand channel pitch (elements), K = kH*kW*Cp, K blocks
  int pix_row;                   // bytes of one halo pixel in a plane: min(pixel bytes, 128)
  int nblk;                      // 128-byte channel blocks per pixel (1 when the pixel is <= 128 B)
  int plane_bytes;               // shared-memory bytes of one (operand, block) plane, 1024-aligned
  int nb;                        // weight ring slots; nb >= num_kb (one N tile): weights resident
  int nhalo;                     // halo buffers in use (<= TL_NHALO)
  int layout;                    // UMMA layout type of the A descriptors
  int Cout, CoutPad, Op, relu;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

// A-operand descriptor of a halo plane: start address filled in per K step
__device__ __forceinline__ uint64_t tile_adesc_base(uint32_t sbo_bytes, int layout) {
  return ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

template <typename T, typename TO, bool SPLIT3, int BN>
__global__ void __launch_bounds__(128 + um_epi(BN))
conv_tile_kernel(const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo,
                 const __grid_constant__ CUtensorMap wmap, const int32_t* __restrict__ tile_ws,
                 const uint32_t* __restrict__ dil_bits, const float* __restrict__ bias,
                 TO* __restrict__ out, const TileGeom g) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int ES = sizeof(T), BK = UM_ROW_BYTES / ES, UK = 32 / ES, KS = BK / UK;
  constexpr int NSPLIT = SPLIT3 ? 2 : 1;
  constexpr int B_BYTES = BN * UM_ROW_BYTES, B_STAGE = NSPLIT * B_BYTES;
  constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  constexpr int EPI = um_epi(BN);
  const int ntl = tile_ws[1];                               // dirty tiles (cb_dilate_compact_tiles)
  const int ntiles_n = g.CoutPad / BN;
  const long long total = (long long)ntl * ntiles_n;
  if ((long long)blockIdx.x >= total) return;               // CTA-uniform, before any barrier / alloc
  if (smem_u32(smem) & 1023u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = g.B * g.TY * g.TXp;
  const int32_t* tiles = tile_ws + 4 + NT;
  const int halo_stage = NSPLIT * g.nblk * g.plane_bytes;
  uint8_t* bring = smem + g.nhalo * halo_stage;
  TileCtrl* ctrl = reinterpret_cast<TileCtrl*>(bring + g.nb * B_STAGE);
  uint2* tab = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(ctrl) + TL_CTRL_BYTES);
  const bool resident = ntiles_n == 1 && g.num_kb <= g.nb;
  const int ph = (g.kH - 1) / 2, pw = (g.kW - 1) / 2;

  // ---- K-step table: per tcgen05.mma (32 bytes of K) the byte offset of its tap inside a halo
  //      plane (+ channel block plane, + channel offset inside the pixel) and, for 16-byte pixels,
  //      the distance to the second tap of the instruction (LBO) ---------------------------------
  {
    const int pixb = g.Cp * ES;
    for (int e = tid; e < g.num_kb * KS; e += blockDim.x) {
      const int k0 = e * UK;
      uint2 v = make_uint2(0xffffffffu, 16u);                // beyond K: all-zero weights, skipped
      if (k0 < g.Kp) {
        const int tap = k0 / g.Cp, ci0 = k0 - tap * g.Cp;
        const int ky = tap / g.kW, kx = tap - ky * g.kW;
        if (pixb == 16) {
          const int t1 = tap + 1, ky1 = t1 / g.kW, kx1 = t1 - ky1 * g.kW;
          v.x = (uint32_t)((ky * g.HWX + kx) * 16);
          // (a second tap beyond the filter meets zero weights; it re-reads the first tap so that no
          //  uninitialised shared memory -- possibly NaN patterns -- enters the product)
          v.y = t1 < g.kH * g.kW ? (uint32_t)(((ky1 * g.HWX + kx1) - (ky * g.HWX + kx)) * 16) : 0u;
        } else {
          const int byte = ci0 * ES;
          const int blk = byte >> 7, within = byte & 127;    // (pixels <= 128 B: blk == 0)
          v.x = (uint32_t)(blk * g.plane_bytes + (ky * g.HWX + kx) * g.pix_row + within);
        }
      }
      tab[e] = v;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < TL_MAXB; ++s) {
      mbar_init(&ctrl->b_full[s], 1);
      mbar_init(&ctrl->b_empty[s], 1);
    }
    for (int b = 0; b < TL_NHALO; ++b) {
      mbar_init(&ctrl->halo_full[b], 1);
      mbar_init(&ctrl->halo_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctrl->tmem_full[b], 1);
      mbar_init(&ctrl->tmem_empty[b], EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&ctrl->tmem_base)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wmap)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&amap_hi)) : "memory");
    if (SPLIT3) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&amap_lo)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  const int tiles_per_img = g.TY * g.TXp;

  if (warp == 0) {
    // =============================== weight tiles (TMA ring) ================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        if (resident && it > 0) break;
        const int nt = (int)(w % ntiles_n);
        for (int kb = 0; kb < g.num_kb; ++kb) {
          if (!resident) mbar_wait(&ctrl->b_empty[stage], phase ^ 1u);
          uint8_t* b_hi = bring + stage * B_STAGE;
          mbar_arrive_expect_tx(&ctrl->b_full[stage], (uint32_t)B_STAGE);
          tma_load_2d(smem_u32(b_hi), &wmap, kb * BK, nt * BN, &ctrl->b_full[stage]);
          if (SPLIT3)
            tma_load_2d(smem_u32(b_hi + B_BYTES), &wmap, kb * BK, g.CoutPad + nt * BN, &ctrl->b_full[stage]);
          if (++stage == (uint32_t)g.nb) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== halo tiles (TMA, tile mode, zero fill) =================
    if (lane == 0) {
      const int BE = g.pix_row / ES;                         // channels per 128-byte block
      const uint32_t box_bytes = (uint32_t)(g.HWX * g.HWY * g.pix_row);
      int it = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int tile = __ldg(tiles + (int)(w / ntiles_n));
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        const int ty = r / g.TXp, tx = r - ty * g.TXp;
        const int hb = it % g.nhalo;
        mbar_wait(&ctrl->halo_empty[hb], (uint32_t)(((it / g.nhalo) & 1) ^ 1));
        mbar_arrive_expect_tx(&ctrl->halo_full[hb], (uint32_t)(NSPLIT * g.nblk) * box_bytes);
        const uint32_t dst = smem_u32(smem + hb * halo_stage);
        for (int blk = 0; blk < g.nblk; ++blk) {
          tma_load_4d(dst + blk * g.plane_bytes, &amap_hi, blk * BE, tx * TL_W - pw, ty * TL_H - ph, b,
                      &ctrl->halo_full[hb]);
          if (SPLIT3)
            tma_load_4d(dst + (g.nblk + blk) * g.plane_bytes, &amap_lo, blk * BE, tx * TL_W - pw,
                        ty * TL_H - ph, b, &ctrl->halo_full[hb]);
        }
      }
    }
  } else if (warp == 2) {
    // =============================== MMA issuer ==============================================
    if (lane == 0) {
      constexpr int KIND = sizeof(T) == 4 ? 0 : 1;
      const uint32_t idesc = umma_idesc(sizeof(T) == 4 ? 2 : (std::is_same<T, __half>::value ? 0 : 1), BN);
      const uint64_t abase = tile_adesc_base((uint32_t)(g.HWX * g.pix_row), g.layout);
      const uint32_t lo_plane = (uint32_t)(g.nblk * g.plane_bytes);
      const uint32_t tab_s = smem_u32(tab);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int hb = it % g.nhalo;
        const uint32_t ab = (uint32_t)it & 1u, aph = ((uint32_t)it >> 1) & 1u;
        mbar_wait(&ctrl->halo_full[hb], (uint32_t)((it / g.nhalo) & 1));
        mbar_wait(&ctrl->tmem_empty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + hb * halo_stage);
        const uint32_t tmem_d = tmem_base + ab * (uint32_t)BN;
        uint32_t acc = 0;
        for (int kb = 0; kb < g.num_kb; ++kb) {
          if (resident) {
            stage = (uint32_t)kb;
            if (it == 0) mbar_wait(&ctrl->b_full[stage], 0u);
          } else {
            mbar_wait(&ctrl->b_full[stage], phase);
          }
          tc_fence_after();
          const uint32_t b_hi = smem_u32(bring + stage * B_STAGE);
          const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t off, lbo;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                         : "=r"(off), "=r"(lbo)
                         : "r"(tab_s + (uint32_t)((kb * KS + ks) * 8)));
            if (off == 0xffffffffu) continue;
            const uint64_t ad = abase | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16);
            const uint64_t ad_hi = ad | (uint64_t)(((a_hi + off) & 0x3FFFFu) >> 4);
            const uint32_t adv = (uint32_t)(ks * 32);
            if (SPLIT3) {
              const uint64_t ad_lo = ad | (uint64_t)(((a_hi + lo_plane + off) & 0x3FFFFu) >> 4);
              umma<KIND>(tmem_d, ad_lo, umma_desc(b_hi + adv), idesc, acc);
              umma<KIND>(tmem_d, ad_hi, umma_desc(b_lo + adv), idesc, 1u);
              umma<KIND>(tmem_d, ad_hi, umma_desc(b_hi + adv), idesc, 1u);
            } else {
              umma<KIND>(tmem_d, ad_hi, umma_desc(b_hi + adv), idesc, acc);
            }
            acc = 1u;
          }
          if (!resident) {
            umma_commit(&ctrl->b_empty[stage]);
            if (++stage == (uint32_t)g.nb) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(&ctrl->tmem_full[ab]);                   // accumulator complete
        umma_commit(&ctrl->halo_empty[hb]);                  // halo buffer consumed
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue: TMEM -> bias / ReLU -> changed rows only ======
    const int q = warp & 3;                                  // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const int rty = row >> 3, rtx = row & 7;
    constexpr int OVEC = 16 / (int)sizeof(TO);
    constexpr int NGROUP = EPI / 128;
    constexpr int COLS = (BN / NGROUP) < 16 ? 16 : (BN / NGROUP);
    const int cbeg = ((warp - 4) >> 2) * COLS;
    int it = 0;
    for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int tile = __ldg(tiles + (int)(w / ntiles_n));
      const int nt = (int)(w % ntiles_n);
      const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
      const int ty = r / g.TXp, tx = r - ty * g.TXp;
      const int y = ty * TL_H + rty, x = tx * TL_W + rtx;
      bool on = false;
      if (y < g.H && x < g.W)
        on = (__ldg(dil_bits + ((long long)b * g.H + y) * g.Wd + (x >> 5)) >> (x & 31)) & 1u;
      const uint32_t ab = (uint32_t)it & 1u, aph = ((uint32_t)it >> 1) & 1u;
      mbar_wait(&ctrl->tmem_full[ab], aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ab * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      TO* orow = out + (((long long)b * g.H + (on ? y : 0)) * g.W + (on ? x : 0)) * g.Op;
      if (__any_sync(0xffffffffu, on)) {
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + COLS && c0 < BN; c0 += 16) {
          uint32_t acc[16];
          tmem_ld16(trow + (uint32_t)c0, acc);
          tmem_ld_wait();
          const int co0 = nt * BN + c0;
          if (on && co0 < g.Cout) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int co = co0 + i;
              float t = __uint_as_float(acc[i]) + (co < g.Cout ? __ldg(bias + co) : 0.f);
              if (g.relu && t <= 0.f) t = 0.f;
              f[i] = t;
            }
            if (co0 + 16 <= g.Cout && (g.Op % OVEC) == 0) {
              if (sizeof(TO) == 4) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + co0 + i) =
                      make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                  TO h[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) h[e] = from_float<TO>(f[i + e]);
                  *reinterpret_cast<uint4*>(orow + co0 + i) = *reinterpret_cast<uint4*>(h);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (co0 + i < g.Cout) orow[co0 + i] = from_float<TO>(f[i]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctrl->tmem_empty[ab]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
struct TilePlan {
  bool ok;
  TileGeom g;
  int smem_bytes, occ;
  long long mma_clk_per_tile;    // rough tensor-pipe time of one tile (policy: cheap tiles only)
};

// Can (and should) this layer run on the tile path?  es = operand element bytes, Cp = operand channel
// pitch, bn = N tile.
inline TilePlan tile_plan(int es, bool split3, int bn, int Cp, int B, int H, int W, int Cout,
                          int CoutPad, int Op, int kH, int kW, int relu) {
  TilePlan p;
  p.ok = false;
  TileGeom& g = p.g;
  const int pixb = Cp * es;
  if (!(pixb == 16 || pixb == 32 || pixb == 64 || (pixb % 128) == 0)) return p;
  if (kH * kW == 1 || !(kH & 1) || !(kW & 1)) return p;      // 1x1: nothing to reuse
  const int nsplit = split3 ? 2 : 1;
  const int bk = UM_ROW_BYTES / es;
  g.B = B; g.H = H; g.W = W; g.Wd = (W + 31) / 32;
  g.TY = tile_grid_y(H); g.TXp = tile_grid_xp(W);
  g.kH = kH; g.kW = kW; g.HWX = TL_W + kW - 1; g.HWY = TL_H + kH - 1;
  if (g.HWX > 256 || g.HWY > 256) return p;
  g.Cp = Cp; g.Kp = kH * kW * Cp;
  g.num_kb = (g.Kp + bk - 1) / bk;
  g.pix_row = pixb < 128 ? pixb : 128;
  g.nblk = pixb > 128 ? pixb / 128 : 1;
  g.plane_bytes = (g.HWX * g.HWY * g.pix_row + 1023) / 1024 * 1024;
  g.layout = g.pix_row == 16 ? 0 : g.pix_row == 32 ? 6 : g.pix_row == 64 ? 4 : 2;
  g.Cout = Cout; g.CoutPad = CoutPad; g.Op = Op; g.relu = relu;
  if (g.num_kb * 4 > TL_MAXTAB) return p;
  if ((long long)g.HWX * g.pix_row >= (1 << 18)) return p;
  const int b_stage = nsplit * bn * UM_ROW_BYTES;
  const int tmem_cols = 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512;
  const int budget2 = 112 * 1024, budget1 = 224 * 1024;
  const bool one_ntile = CoutPad == bn;
  int nb = 0, occ = 1, fixed = 0;
  g.nhalo = 0;
  for (int nh = TL_NHALO; nh >= 1 && !g.nhalo; --nh) {       // fewer halo buffers when smem is tight
    fixed = nh * nsplit * g.nblk * g.plane_bytes + TL_CTRL_BYTES + g.num_kb * 4 * 8;
    if (nh == TL_NHALO && one_ntile && g.num_kb <= TL_MAXB && fixed + g.num_kb * b_stage <= budget2 &&
        2 * tmem_cols <= 512) {
      nb = g.num_kb; occ = 2; g.nhalo = nh;                  // resident weights, two CTAs per SM
    } else if (nh == TL_NHALO && budget2 > fixed && (budget2 - fixed) / b_stage >= 4 && 2 * tmem_cols <= 512) {
      nb = (budget2 - fixed) / b_stage; occ = 2; g.nhalo = nh;
    } else if (budget1 > fixed && (budget1 - fixed) / b_stage >= (nh == 1 ? 2 : 3)) {
      nb = (budget1 - fixed) / b_stage; occ = 1; g.nhalo = nh;
      if (one_ntile && g.num_kb <= TL_MAXB && nb >= g.num_kb) nb = g.num_kb;
    }
  }
  if (!g.nhalo) return p;
  if (nb > TL_MAXB) nb = TL_MAXB;
  static const int force_occ = [] {
    const char* e = getenv("CBINFER_TILE_OCC");              // tuning knob: 1 = one CTA per SM
    return e ? atoi(e) : 0;
  }();
  if (force_occ == 1 && occ == 2) {
    occ = 1;
    nb = (budget1 - fixed) / b_stage;
    if (nb > TL_MAXB) nb = TL_MAXB;
    if (one_ntile && g.num_kb <= nb) nb = g.num_kb;
  }
  g.nb = nb;
  p.occ = occ;
  p.smem_bytes = fixed + nb * b_stage;
  const int uk = 32 / es;
  const long long ksteps = (g.Kp + uk - 1) / uk;
  const long long per = bn / 2 < 32 ? 32 : bn / 2;           // cycles per instruction (N/2, smem floor ~32)
  p.mma_clk_per_tile = ksteps * (split3 ? 3 : 1) * per * (CoutPad / bn);
  p.ok = true;
  return p;
}

template <typename T>
inline CUtensorMapDataType tmap_dtype() {
  return sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
         : std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                          : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

template <typename T, typename TO, bool SPLIT3, int BN>
int launch_conv_tile(cudaStream_t s, const void* state, const void* state_lo, const int32_t* tile_ws,
                     const uint32_t* dil_bits, const void* packed, const float* bias, void* out,
                     const TilePlan& plan) {
  const TileGeom& g = plan.g;
  auto enc = tensor_map_encoder();
  if (!enc) return fail(3, "conv_update_tiled: cuTensorMapEncodeTiled unavailable");
  constexpr int ES = sizeof(T), BK = UM_ROW_BYTES / ES, NSPLIT = SPLIT3 ? 2 : 1;
  const int KpPad = g.num_kb * BK;
  alignas(64) CUtensorMap wmap, amap[2];
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)KpPad, (cuuint64_t)(NSPLIT * g.CoutPad)};
    const cuuint64_t gstr[1] = {(cuuint64_t)KpPad * ES};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&wmap, tmap_dtype<T>(), 2, const_cast<void*>(packed), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(3, "conv_update_tiled: weight tensor map failed (%d)", (int)r);
  }
  const CUtensorMapSwizzle sw = g.pix_row == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : g.pix_row == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : g.pix_row == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  for (int pl = 0; pl < NSPLIT; ++pl) {
    const void* base = pl ? state_lo : state;
    const cuuint64_t pixb = (cuuint64_t)g.Cp * ES;
    const cuuint64_t gdim[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
    const cuuint64_t gstr[3] = {pixb, pixb * g.W, pixb * g.W * g.H};
    const cuuint32_t box[4] = {(cuuint32_t)(g.pix_row / ES), (cuuint32_t)g.HWX, (cuuint32_t)g.HWY, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&amap[pl], tmap_dtype<T>(), 4, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(3, "conv_update_tiled: state tensor map failed (%d)", (int)r);
  }
  if (!SPLIT3) amap[1] = amap[0];
  auto kern = conv_tile_kernel<T, TO, SPLIT3, BN>;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail(3, "conv_update_tiled: cannot reserve shared memory");
    attr_dev = dev;
  }
  const long long max_items = (long long)g.B * g.TY * tile_grid_x(g.W) * (g.CoutPad / BN);
  long long grid = (long long)sm_count() * plan.occ;
  if (grid > max_items) grid = max_items;
  if (grid < 1) grid = 1;
  cb::launch_pdl(kern, dim3((unsigned)grid), dim3(128 + um_epi(BN)), (size_t)plan.smem_bytes, s, amap[0],
                 amap[1], wmap, tile_ws, dil_bits, bias, (TO*)out, g);
  CB_CHECK_LAUNCH("conv_update_tiled");
  return 0;
}

// policy threshold: tiles whose tensor work is at most this many cycles (denser layers keep the
// index-list kernel, whose M tiles hold changed pixels only)
inline long long tile_clk_limit() {
  static const long long v = [] {
    const char* e = getenv("CBINFER_TILE_CLK");
    return e ? atoll(e) : 12000ll;
  }();
  return v;
}

inline int umma_tile_plan(TilePlan& plan, int dtype, int gemm, int Cp, int B, int H, int W, int Cout,
                          int Op, int kH, int kW, int relu) {
  if (!(gemm == CB_GEMM_TC || gemm == CB_GEMM_TC_3X || gemm == CB_GEMM_TC_BF16X3)) { plan.ok = false; return 0; }
  const int bn = umma_bn(gemm, Cout), CoutPad = umma_cout_pad(gemm, Cout);
  plan = tile_plan(umma_operand_es(dtype, gemm), umma_is_split(dtype, gemm), bn, Cp, B, H, W, Cout,
                   CoutPad, Op, kH, kW, relu);
  return 0;
}

inline int umma_conv_update_tiled(cudaStream_t s, int dtype, int gemm, const void* state,
                                  const void* state_lo, int Cp, const int32_t* tile_ws,
                                  const uint32_t* dil_bits, const void* packed, const float* bias,
                                  void* out, int Op, int B, int H, int W, int Cout, int kH, int kW,
                                  int relu) {
  TilePlan plan;
  umma_tile_plan(plan, dtype, gemm, Cp, B, H, W, Cout, Op, kH, kW, relu);
  CB_CHECK_ARG(plan.ok, "conv_update_tiled: layer shape not supported by the tile path");
  const bool split3 = gemm == CB_GEMM_TC_3X && dtype == CB_F32;
  const bool bf16x3 = gemm == CB_GEMM_TC_BF16X3;
  CB_CHECK_ARG(!(split3 || bf16x3) || state_lo, "conv_update_tiled: the 3x modes need state_lo");
  CB_CHECK_ARG(((uintptr_t)state % 16) == 0 && ((uintptr_t)packed % 128) == 0,
               "conv_update_tiled: state must be 16-byte and packed weights 128-byte aligned");
  const int bn = umma_bn(gemm, Cout);
#define CB_TBN(T_, TO_, S3_)                                                                       \
  switch (bn) {                                                                                    \
    case 16: return launch_conv_tile<T_, TO_, S3_, 16>(s, state, state_lo, tile_ws, dil_bits, packed, bias, out, plan);   \
    case 32: return launch_conv_tile<T_, TO_, S3_, 32>(s, state, state_lo, tile_ws, dil_bits, packed, bias, out, plan);   \
    case 64: return launch_conv_tile<T_, TO_, S3_, 64>(s, state, state_lo, tile_ws, dil_bits, packed, bias, out, plan);   \
    case 128: return launch_conv_tile<T_, TO_, S3_, 128>(s, state, state_lo, tile_ws, dil_bits, packed, bias, out, plan); \
    case 256: return launch_conv_tile<T_, TO_, S3_, 256>(s, state, state_lo, tile_ws, dil_bits, packed, bias, out, plan); \
    default: return fail(2, "conv_update_tiled: unsupported N tile %d", bn);                      \
  }
  if (bf16x3) { CB_TBN(__nv_bfloat16, float, true) }
  switch (dtype) {
    case CB_F32:
      if (split3) { CB_TBN(float, float, true) } else { CB_TBN(float, float, false) }
    case CB_F16: CB_TBN(__half, __half, false)
    case CB_BF16: CB_TBN(__nv_bfloat16, __nv_bfloat16, false)
    default: return fail(2, "conv_update_tiled: bad dtype %d", dtype);
  }
#undef CB_TBN
}

}  // namespace cb
