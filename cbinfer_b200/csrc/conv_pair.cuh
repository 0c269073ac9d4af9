// conv_pair.cuh -- the gather contraction on CTA PAIRS (tcgen05.mma.cta_group::2): wide layers with many
// changed pixels.
//
// Same job and the same arithmetic as conv_umma.cuh (reference genXMatrix_kernel cbconv2d_cg_backend.cu:138-161
// + the cuBLAS GEMM conv2d_cg.py:342-349 + updateOutput_kernel cbconv2d_cg_backend.cu:175-189): per K block the
// index-list kernel moves 32 KB of gathered state (hi + lo planes of 128 pixels) AND 64 KB of weights (N = 256)
// into every SM, and the 64 -> 256 7x7 layer sits on that L2 -> SM ingest (~62 B/clk per SM, tensor pipe 64 %).
// Here two CTAs on the two SMs of a TPC execute ONE M = 256 instruction: each CTA gathers the 128 pixels of
// its own M tile and loads only ITS HALF of the weight tile (N rows 128 r .. 128 r + 127; probed on the B200:
// tools/umma_pair_probe.cu) -- 64 KB per stage and SM instead of 96, three stages instead of two.
//
//   work item  = (pair of consecutive M tiles, N tile of 256): CTA r of the cluster owns M tile 2 p + r
//   warps 0-7  gather producers (cp.async into the 128B-swizzled K-major tile, as conv_umma.cuh)
//   warp  8    TMA: the CTA's N half of the weight tile (hi and lo planes)
//   warp  9    leader CTA: issues tcgen05.mma.cta_group::2 (M = 256), releases stages / accumulators in BOTH
//              CTAs with multicast commits; peer CTA: relays "my stage is full" to the leader's barrier
//   warps 10.. epilogue of the CTA's own 128 rows (TMEM -> bias / ReLU -> one contiguous channel run per pixel);
//              two accumulators, so the epilogue of one item overlaps the MMAs of the next
// K order and the 3-term split are those of conv_umma.cuh: every output value is the same sum of the same
// products in the same order -- bit-identical results.  No stream-K here: the kernel takes over from
// `sel_lo` M tiles on (enough pairs for every SM), the index-list kernel keeps the small counts.
#pragma once
#include "conv_umma.cuh"

namespace cb {

constexpr int PR_BN = 256, PR_NH = 128;                      // N tile of the pair / rows of it per CTA
// Stage = PR_ROWB bytes of K per row.  128-byte rows (64 bf16, SWIZZLE_128B), three stages.  Measured: 64-byte rows
// (SWIZZLE_64B) with six stages -- same bytes in flight, twice the barrier round trips -- are much slower
// (64 -> 256 7x7 layer, 14k changed pixels: 98 vs 69 us in step).
constexpr int PR_ROWB = 128;                                 // bytes of K per stage row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
constexpr int PR_STAGES = PR_ROWB == 128 ? 3 : 6;
constexpr int PR_EPI = um_epi(PR_BN), PR_THREADS = um_threads(PR_BN);

struct PairCtrl {
  uint64_t full[PR_STAGES], empty[PR_STAGES];
  uint64_t peer_full[PR_STAGES];               // leader only: the peer CTA's stage is full
  uint64_t tmem_full[2], tmem_empty[2];        // per accumulator: MMAs done / drained by MY epilogue
  uint64_t pair_empty[2];                      // leader only: drained by BOTH epilogues
  uint64_t tab_full[2];
  uint32_t tmem_base, pad;
  int pix[2][UM_BM];
  int yx[2][UM_BM];
};
static_assert(sizeof(PairCtrl) <= UM_CTRL_BYTES, "ctrl block too large");

template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (KIND == 0)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs are done
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <typename T, bool SPLIT3>
struct PairCfg {
  static constexpr int ES = sizeof(T), VEC = 16 / ES, BK = PR_ROWB / ES, UK = 32 / ES;
  static constexpr int CPR = PR_ROWB / 16;                     // 16-byte chunks per stage row
  static constexpr int NSPLIT = SPLIT3 ? 2 : 1;
  static constexpr int A_BYTES = UM_BM * PR_ROWB;              // 8 KB
  static constexpr int B_BYTES = PR_NH * PR_ROWB;              // 8 KB: this CTA's half of the N tile
  static constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + B_BYTES);
  static constexpr int TABLE_MAX = 1024;
};

template <typename T, typename TO, bool SPLIT3>
__global__ void __launch_bounds__(PR_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap wmap, const T* __restrict__ state,
                 const T* __restrict__ state_lo, int Cp, const int32_t* __restrict__ idx,
                 const int32_t* __restrict__ count, const float* __restrict__ bias, TO* __restrict__ out, int Op,
                 int H, int W, int Cout, int CoutPad, int kH, int kW, int Kp, int relu, int sel_lo, int sel_hi) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem[];
  using C = PairCfg<T, SPLIT3>;
  constexpr int RPT = UM_BM * C::CPR / UM_PRODUCERS, RSTEP = UM_PRODUCERS / C::CPR;
  const int n = __shfl_sync(0xffffffffu, *count, 0);
  const int mtiles = (n + UM_BM - 1) / UM_BM;
  if (mtiles < sel_lo || mtiles >= sel_hi) return;          // (grid-uniform: before any barrier / alloc)
  const int ntiles = CoutPad / PR_BN;
  const int npairs = (mtiles + 1) / 2;
  const long long total = (long long)npairs * ntiles;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);
  if ((long long)pair0 >= total) return;                    // (cluster-uniform)
  if (smem_u32(smem) & 1023u) __trap();
  PairCtrl* ctrl = reinterpret_cast<PairCtrl*>(smem + PR_STAGES * C::STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int num_kb = (Kp + C::BK - 1) / C::BK;
  constexpr int TMA_WARP = UM_PRODUCERS / 32, MMA_WARP = TMA_WARP + 1, EPI_WARP0 = MMA_WARP + 1;
  const int ph = (kH - 1) / 2, pw = (kW - 1) / 2;
  int2* ktab = reinterpret_cast<int2*>(reinterpret_cast<uint8_t*>(ctrl) + UM_CTRL_BYTES);
  const uint32_t ktab_s = smem_u32(ktab);
  for (int q = tid; q < num_kb * C::CPR; q += PR_THREADS) {  // (host: num_kb * CPR <= TABLE_MAX)
    const int k = q * C::VEC;
    int2 e = make_int2(0, (int)0x80008000u);
    if (k < Kp) {
      const int tap = k / Cp, ci = k - tap * Cp;
      const int ky = tap / kW, kx = tap - ky * kW;
      e.x = ((ky - ph) * W + (kx - pw)) * Cp + ci;
      e.y = (int)(((unsigned)(ky - ph) << 16) | ((unsigned)(kx - pw) & 0xffffu));
    }
    ktab[q] = e;
  }
  if (tid == 0) {
    for (int s = 0; s < PR_STAGES; ++s) {
      mbar_init(&ctrl->full[s], UM_PRODUCERS + 1);
      mbar_init(&ctrl->empty[s], 1);
      mbar_init(&ctrl->peer_full[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctrl->tmem_full[b], 1);
      mbar_init(&ctrl->tmem_empty[b], PR_EPI);
      mbar_init(&ctrl->pair_empty[b], 2 * PR_EPI);
      mbar_init(&ctrl->tab_full[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&ctrl->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == TMA_WARP && lane == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wmap)) : "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // the peer's barriers and TMEM exist
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  const int P = H * W;

  if (warp < TMA_WARP) {
    // =============================== gather producers (my M tile) =============================
    uint32_t stage = 0, phase = 0, sidx = 0;
    const int c = tid & (C::CPR - 1), r0 = tid / C::CPR;
    uint32_t soff[RPT];
#pragma unroll
    for (int it = 0; it < RPT; ++it) {     // 8-row groups; 16-byte chunk ^ row (SWIZZLE_128B) / ^ (row / 2 & 3) (SWIZZLE_64B)
      const int r = r0 + RSTEP * it;
      const int x = PR_ROWB == 128 ? (r & 7) : ((r >> 1) & 3);
      soff[it] = (uint32_t)((r >> 3) * (8 * PR_ROWB) + (r & 7) * PR_ROWB + ((c ^ x) << 4));
    }
    const long long lo_delta = SPLIT3 ? (state_lo - state) : 0;
    for (long long w = pair0; w < total; w += pair_step) {
      const int mt = 2 * (int)(w / ntiles) + (int)rank;
      const uint32_t buf = sidx & 1u, use = sidx >> 1;
      ++sidx;
      if (tid < UM_BM) {
        mbar_wait(&ctrl->tmem_empty[buf], (use & 1u) ^ 1u); // my epilogue has drained this table / accumulator
        const int j = mt * UM_BM + tid;
        int pix = -1, yx = 0;
        if (j < n) {
          pix = __ldg(idx + j);
          const int p = pix % P;
          const int yy = p / W;
          yx = (yy << 16) | (p - yy * W);
        }
        ctrl->pix[buf][tid] = pix;
        ctrl->yx[buf][tid] = yx;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(UM_PRODUCERS) : "memory");
      if (tid == 0) mbar_arrive(&ctrl->tab_full[buf]);
      const T* rbase[RPT];
      int ry[RPT], rx[RPT];
#pragma unroll
      for (int it = 0; it < RPT; ++it) {
        const int pix = ctrl->pix[buf][r0 + RSTEP * it], yx = ctrl->yx[buf][r0 + RSTEP * it];
        rbase[it] = state + (long long)(pix < 0 ? 0 : pix) * Cp;
        ry[it] = pix < 0 ? -0x40000000 : (yx >> 16);
        rx[it] = yx & 0xffff;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&ctrl->empty[stage], phase ^ 1u);
        const uint32_t a_hi = smem_u32(smem + stage * C::STAGE_BYTES);
        int2 e;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(ktab_s + (uint32_t)((kb * C::CPR + c) * 8)));
        const long long koff = e.x;
        const int dy = e.y >> 16, dx = (int)(short)(e.y & 0xffff);
#pragma unroll
        for (int it = 0; it < RPT; ++it) {
          const bool ok = (unsigned)(ry[it] + dy) < (unsigned)H && (unsigned)(rx[it] + dx) < (unsigned)W;
          const T* src = ok ? rbase[it] + koff : state;
          cp_async16(a_hi + soff[it], src, ok ? 16u : 0u);
          if (SPLIT3) cp_async16(a_hi + C::A_BYTES + soff[it], src + lo_delta, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(&ctrl->full[stage]);
        if (++stage == PR_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == TMA_WARP) {
    // =============================== my half of the weight tiles ==============================
    const bool leader = elect_one();
    uint32_t stage = 0, phase = 0;
    for (long long w = pair0; w < total; w += pair_step) {
      const int nt = (int)(w % ntiles);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&ctrl->empty[stage], phase ^ 1u);
        if (leader) {
          const uint32_t b_hi = smem_u32(smem + stage * C::STAGE_BYTES + C::NSPLIT * C::A_BYTES);
          mbar_arrive_expect_tx(&ctrl->full[stage], (uint32_t)(C::NSPLIT * C::B_BYTES));
          tma_load_2d(b_hi, &wmap, kb * C::BK, nt * PR_BN + (int)rank * PR_NH, &ctrl->full[stage]);
          if (SPLIT3)
            tma_load_2d(b_hi + C::B_BYTES, &wmap, kb * C::BK, CoutPad + nt * PR_BN + (int)rank * PR_NH, &ctrl->full[stage]);
        }
        __syncwarp();
        if (++stage == PR_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == MMA_WARP) {
    uint32_t stage = 0, phase = 0, sidx = 0;
    if (rank != 0) {
      // =============================== peer: relay "stage full" to the leader ==================
      const bool one = elect_one();
      for (long long w = pair0; w < total; w += pair_step)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&ctrl->full[stage], phase);
          fence_proxy_async_smem();                          // my cp.async writes -> async proxy (the pair's UMMA)
          if (one) mbar_arrive_remote(map_to_rank(smem_u32(&ctrl->peer_full[stage]), 0));
          __syncwarp();
          if (++stage == PR_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else {
      // =============================== leader: MMA issuer (M = 256) ============================
      constexpr int KIND = sizeof(T) == 4 ? 0 : 1;
      const uint32_t fmt = sizeof(T) == 4 ? 2u : (std::is_same<T, __half>::value ? 0u : 1u);
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(PR_BN >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      const bool one = elect_one();
      // K-major swizzled descriptor: LBO 1, SBO = 8 rows x PR_ROWB, version 1, layout 2 (128 B) / 4 (64 B)
      const uint32_t d_hi32 = (uint32_t)((((uint64_t)((8 * PR_ROWB) >> 4) << 32) | (1ull << 46) |
                                          ((PR_ROWB == 128 ? 2ull : 4ull) << 61)) >> 32);
      for (long long w = pair0; w < total; w += pair_step) {
        const uint32_t buf = sidx & 1u, use = sidx >> 1;
        ++sidx;
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)PR_BN;
        mbar_wait_cluster(&ctrl->pair_empty[buf], (use & 1u) ^ 1u);   // both epilogues drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&ctrl->full[stage], phase);
          mbar_wait_cluster(&ctrl->peer_full[stage], phase);
          fence_proxy_async_smem();
          tc_fence_after();
          const uint32_t a_hi = ((smem_u32(smem + stage * C::STAGE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t a_lo = a_hi + (uint32_t)(C::A_BYTES >> 4);
          const uint32_t b_hi = a_hi + (uint32_t)((C::NSPLIT * C::A_BYTES) >> 4);
          const uint32_t b_lo = b_hi + (uint32_t)(C::B_BYTES >> 4);
          if (one) {
#pragma unroll
            for (int ks = 0; ks < C::BK / C::UK; ++ks) {
              const uint32_t adv = (uint32_t)(ks * 2);
              const uint32_t first = (kb || ks) ? 1u : 0u;
              const uint64_t dA = ((uint64_t)d_hi32 << 32) | (a_hi + adv);
              const uint64_t dB = ((uint64_t)d_hi32 << 32) | (b_hi + adv);
              if (SPLIT3) {
                const uint64_t dAl = ((uint64_t)d_hi32 << 32) | (a_lo + adv);
                const uint64_t dBl = ((uint64_t)d_hi32 << 32) | (b_lo + adv);
                umma_pair<KIND>(tmem_d, dAl, dB, idesc, first);
                umma_pair<KIND>(tmem_d, dA, dBl, idesc, 1u);
                umma_pair<KIND>(tmem_d, dA, dB, idesc, 1u);
              } else {
                umma_pair<KIND>(tmem_d, dA, dB, idesc, first);
              }
            }
            umma_commit_pair(&ctrl->empty[stage]);           // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == PR_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (one) umma_commit_pair(&ctrl->tmem_full[buf]);    // accumulator complete, in both CTAs
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue of my 128 rows ==================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    constexpr int OVEC = 16 / (int)sizeof(TO);
    constexpr int COLS = PR_BN / (PR_EPI / 128);
    const int cbeg = ((warp - EPI_WARP0) >> 2) * COLS;
    const uint32_t pe_addr = map_to_rank(smem_u32(&ctrl->pair_empty[0]), 0);
    uint32_t sidx = 0;
    for (long long w = pair0; w < total; w += pair_step) {
      const int nt = (int)(w % ntiles);
      const uint32_t buf = sidx & 1u, use = sidx >> 1;
      ++sidx;
      mbar_wait(&ctrl->tab_full[buf], use & 1u);
      mbar_wait(&ctrl->tmem_full[buf], use & 1u);
      tc_fence_after();
      const int pix = ctrl->pix[buf][row];
      const uint32_t trow = tmem_base + buf * (uint32_t)PR_BN + ((uint32_t)(q * 32) << 16);
      TO* orow = out + (long long)(pix < 0 ? 0 : pix) * Op;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + COLS; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(trow + (uint32_t)c0, acc);
        tmem_ld_wait();
        const int co0 = nt * PR_BN + c0;
        if (pix >= 0 && co0 < Cout) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = co0 + i;
            float t = __uint_as_float(acc[i]) + (co < Cout ? __ldg(bias + co) : 0.f);
            if ((relu & 1) && t <= 0.f) t = 0.f;
            f[i] = t;
          }
          if (co0 + 16 <= Cout && (Op % OVEC) == 0) {
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
              if (co0 + i < Cout) orow[co0 + i] = from_float<TO>(f[i]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctrl->tmem_empty[buf]);                   // my gather warps may reuse the row table
      mbar_arrive_remote(pe_addr + buf * (uint32_t)sizeof(uint64_t));   // the leader may reuse the accumulator
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // nobody leaves while the pair's MMAs may read its smem
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// can the pair kernel take this layer?  (one-byte answers only: the count decides on the device)
inline bool pair_supported(int es, int Cp, int CoutPad, int kH, int kW) {
  const int bk = UM_ROW_BYTES / es;                          // (K padding of the packed weights: 64-element blocks)
  const int num_kb = (kH * kW * Cp + bk - 1) / bk;
  return es == 2 && (CoutPad % PR_BN) == 0 && Cp * es >= 16 && !(es == 2 && Cp == 4) && num_kb * 8 <= 1024 && num_kb >= 8;
}

template <typename T, typename TO, bool SPLIT3>
int launch_conv_pair(cudaStream_t s, const void* state, const void* state_lo, int Cp, const int32_t* idx,
                     const int32_t* count, const void* packed, const float* bias, void* out, int Op, int H, int W,
                     int Cout, int CoutPad, int kH, int kW, int relu, int sel_lo, int sel_hi) {
  using C = PairCfg<T, SPLIT3>;
  const int Kp = kH * kW * Cp;
  const int KpPad = umma_kp_pad_es((int)sizeof(T), Cp, kH, kW);
  auto enc = tensor_map_encoder();
  if (!enc) return fail(3, "conv_update: cuTensorMapEncodeTiled unavailable");
  alignas(64) CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)KpPad, (cuuint64_t)(C::NSPLIT * CoutPad)};
  const cuuint64_t gstr[1] = {(cuuint64_t)KpPad * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)C::BK, (cuuint32_t)PR_NH};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                  : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUresult r = enc(&map, dt, 2, const_cast<void*>(packed), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         PR_ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(3, "conv_update(pair): cuTensorMapEncodeTiled failed (%d)", (int)r);
  auto kern = conv_pair_kernel<T, TO, SPLIT3>;
  const int num_kb = KpPad / C::BK;
  const int smem_bytes = PR_STAGES * C::STAGE_BYTES + UM_CTRL_BYTES + num_kb * C::CPR * 8;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail(3, "conv_update(pair): cannot reserve shared memory");
    attr_dev = dev;
  }
  CB_CHECK_ARG(smem_bytes <= 227 * 1024, "conv_update(pair): %d bytes of shared memory", smem_bytes);
  const unsigned grid = (unsigned)(sm_count() / 2 * 2);      // one CTA per SM, in pairs
  cb::launch_cluster(kern, grid, PR_THREADS, (size_t)smem_bytes, s, 2u, 0, map, (const T*)state, (const T*)state_lo, Cp,
                     idx, count, bias, (TO*)out, Op, H, W, Cout, CoutPad, kH, kW, Kp, relu, sel_lo, sel_hi);
  CB_CHECK_LAUNCH("conv_update(pair)");
  return 0;
}

}  // namespace cb
