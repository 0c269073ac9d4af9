// cb_abi.cu -- extern "C" entry points of libcbinfer_sm100.so (see include/cbinfer_b200.h).
// sm_100a only: no other architecture, no CPU path, no dispatch to other backends.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "cb_common.cuh"
#include "compact.cuh"
#include "conv_simt.cuh"
#include "conv_umma.cuh"
#include "conv_pair.cuh"
#include "conv_tile.cuh"
#include "detect.cuh"
#include "ingest.cuh"
#include "fg.cuh"
#include "pool.cuh"
#include "staged.cuh"
#include "tail.cuh"

namespace cb {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    // Off by default: a dependent grid that becomes resident early may see stale L1 / read-only
    // (ld.global.nc) lines of buffers its predecessor is still writing (observed as wrong results
    // inside CUDA graphs); measured gain was only ~2 %.  CBINFER_PDL=1 enables it for experiments.
    const char* e = getenv("CBINFER_PDL");
    cached = (e && e[0] == '1') ? 1 : 0;
  }
  return cached != 0;
}

static inline unsigned grid_for(long long total, int threads, int per_sm = 8) {
  long long blocks = (total + threads - 1) / threads;
  long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace cb

using namespace cb;
static_assert(cb::kDtRows == cb::TL_H && cb::TL_W == 8, "dilate_tiles_kernel lists 8 x 16 tiles, four per bitmap word");

#define CB_DISPATCH_DTYPE(dtype, ...)                                               \
  switch (dtype) {                                                                  \
    case CB_F32: { using T = float; constexpr int VEC = 4; (void)VEC; __VA_ARGS__; } break;           \
    case CB_F16: { using T = __half; constexpr int VEC = 8; (void)VEC; __VA_ARGS__; } break;          \
    case CB_BF16: { using T = __nv_bfloat16; constexpr int VEC = 8; (void)VEC; __VA_ARGS__; } break;  \
    default: return cb::fail(2, "bad dtype %d", dtype);                             \
  }

extern "C" {

int cb_version(void) { return 100; }

const char* cb_last_error(void) { return cb::g_err; }

int cb_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cb::fail(3, "no CUDA device");
  cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev);
  return 0;
}

int cb_bitmap_row_words(int W) { return (W + 31) / 32; }
size_t cb_bitmap_words(int B, int H, int W) { return (size_t)B * H * ((W + 31) / 32); }
size_t cb_compact_ws_bytes(int B, int H, int W) { return cb::compact_ws_bytes(cb_bitmap_words(B, H, W)); }
int cb_plane_pitch16(int C) { return cb::pitch16_of(C); }
int cb_channel_pitch(int dtype, int C) {
  const int v = 16 / cb::esize(dtype);
  return (C + v - 1) / v * v;
}

int cb_change_detect(void* stream, int dtype, const void* x, long long x_sb, long long x_sc,
                     long long x_sy, long long x_sx, void* state, long long s_sb, long long s_sc,
                     long long s_sy, long long s_sx, int aux_mode, void* aux_hi, void* aux_lo,
                     uint32_t* raw_bits, int B, int C, int H, int W, float threshold,
                     int update_mode) {
  CB_CHECK_ARG(x && state && raw_bits, "change_detect: null pointer");
  CB_CHECK_ARG(B >= 0 && C > 0 && H >= 0 && W >= 0, "change_detect: bad shape");
  CB_DISPATCH_DTYPE(dtype, return (launch_detect<T, VEC>((cudaStream_t)stream, x, x_sb, x_sc, x_sy,
                                                        x_sx, state, s_sb, s_sc, s_sy, s_sx, aux_mode,
                                                        aux_hi, aux_lo, raw_bits, B, C, H, W, threshold,
                                                        update_mode)));
  return 0;
}

int cb_change_detect_u8(void* stream, const uint8_t* x, long long x_sb, long long x_sc,
                        long long x_sy, long long x_sx, float* state, long long s_sb, long long s_sc,
                        long long s_sy, long long s_sx, int aux_mode, void* aux_hi, void* aux_lo,
                        uint32_t* raw_bits, int B, int C, int H, int W, float divisor, float bias,
                        float threshold, int update_mode) {
  CB_CHECK_ARG(x && state && raw_bits, "change_detect_u8: null pointer");
  CB_CHECK_ARG(B >= 0 && C > 0 && H >= 0 && W >= 0, "change_detect_u8: bad shape");
  return launch_detect_u8((cudaStream_t)stream, x, x_sb, x_sc, x_sy, x_sx, state, s_sb, s_sc, s_sy,
                          s_sx, aux_mode, aux_hi, aux_lo, raw_bits, B, C, H, W, divisor, bias,
                          threshold, update_mode);
}

static bool compact_coop() {
  static const bool v = [] {
    const char* e = getenv("CBINFER_COMPACT_COOP");          // tuning knob: 0 = per-thread expansion
    return !(e && e[0] == '0');
  }();
  return v;
}

size_t cb_tile_ws_bytes(int B, int H, int W) { return cb::tile_ws_words(B, H, W) * sizeof(int32_t); }

static int dilate_compact_impl(void* stream, const uint32_t* raw_bits, uint32_t* dil_bits, int8_t* dil_map,
                      int32_t* idx, int32_t* count, void* ws, void* tile_ws, int B, int H, int W, int kHHalf,
                      int kWHalf, int clear_raw, int no_list = 0, const cb::PrefetchHints* hints = nullptr) {
  CB_CHECK_ARG(raw_bits && (idx || no_list) && count && ws, "dilate_compact: null pointer");
  cb::PrefetchHints ph;
  memset(&ph, 0, sizeof(ph));
  if (hints) ph = *hints;
  CB_CHECK_ARG(raw_bits != dil_bits, "dilate_compact: dil_bits must not alias raw_bits");
  CB_CHECK_ARG(kHHalf >= 0 && kWHalf >= 0 && kWHalf <= 31, "dilate_compact: kWHalf must be <= 31");
  CB_CHECK_ARG((long long)B * H * W < (1ll << 31), "dilate_compact: more than 2^31 pixels");
  const long long nwords = (long long)cb_bitmap_words(B, H, W);
  cudaStream_t s = (cudaStream_t)stream;
  if (nwords == 0) {
    cudaMemsetAsync(count, 0, sizeof(int32_t), s);
    CB_CHECK_LAUNCH("dilate_compact(memset)");
    return 0;
  }
  CB_CHECK_ARG(nwords < (1ll << 31), "dilate_compact: bitmap too large");
  const int ntiles = (int)((nwords + kCompactTile - 1) / kCompactTile);
  cb::launch_pdl(ph.n ? dilate_compact_kernel<true> : dilate_compact_kernel<false>, ntiles, kCompactThreads, 0, s,
                                                          raw_bits, dil_bits, dil_map, idx, count,
                                                          ws, B, H, W, (W + 31) / 32, kHHalf,
                                                          kWHalf, (int)nwords, ntiles, 0, 0,
                                                          clear_raw ? const_cast<uint32_t*>(raw_bits) : nullptr,
                                                          (int32_t*)tile_ws, cb::tile_grid_y(H), cb::tile_grid_xp(W),
                                                          (int)compact_coop(), no_list, ph);
  CB_CHECK_LAUNCH("dilate_compact");
  return 0;
}

int cb_dilate_compact_hinted(void* stream, const uint32_t* raw_bits, uint32_t* dil_bits, int8_t* dil_map,
                             int32_t* idx, int32_t* count, void* ws, void* tile_ws, int B, int H, int W,
                             int kHHalf, int kWHalf, int clear_raw, int no_list, int n_hints,
                             const void* const* hint_base, const int* hint_row_bytes, const int* hint_shift,
                             const int* hint_h, const int* hint_w) {
  CB_CHECK_ARG(n_hints >= 0 && n_hints <= cb::kMaxHints, "dilate_compact_hinted: at most %d hints", cb::kMaxHints);
  CB_CHECK_ARG(n_hints == 0 || (hint_base && hint_row_bytes && hint_shift && hint_h && hint_w),
               "dilate_compact_hinted: null hint array");
  CB_CHECK_ARG(!no_list || (tile_ws && dil_bits), "dilate_compact_hinted: no_list needs tile_ws and dil_bits");
  cb::PrefetchHints ph;
  memset(&ph, 0, sizeof(ph));
  for (int i = 0; i < n_hints; ++i) {
    CB_CHECK_ARG(hint_shift[i] == 0 || hint_shift[i] == 1, "dilate_compact_hinted: shift must be 0 or 1");
    CB_CHECK_ARG(hint_base[i] && ((uintptr_t)hint_base[i] % 16) == 0 && hint_row_bytes[i] > 0 &&
                     (hint_row_bytes[i] % 16) == 0 && hint_h[i] > 0 && hint_w[i] > 0,
                 "dilate_compact_hinted: hint rows must be 16-byte aligned multiples of 16 bytes");
    // (a hint never reaches beyond its map: rows are clipped to hint_h x hint_w in the kernel)
    CB_CHECK_ARG((long long)hint_row_bytes[i] * 32 < (1ll << 31), "dilate_compact_hinted: row too large");
    ph.base[ph.n] = (const char*)hint_base[i];
    ph.row_bytes[ph.n] = hint_row_bytes[i];
    ph.shift[ph.n] = hint_shift[i];
    ph.tH[ph.n] = hint_h[i];
    ph.tW[ph.n] = hint_w[i];
    ++ph.n;
  }
  if (tile_ws && (long long)cb_bitmap_words(B, H, W) == 0)
    cudaMemsetAsync((int32_t*)tile_ws + 1, 0, sizeof(int32_t), (cudaStream_t)stream);
  return dilate_compact_impl(stream, raw_bits, dil_bits, dil_map, idx, count, ws, tile_ws, B, H, W, kHHalf,
                             kWHalf, clear_raw, no_list, &ph);
}

int cb_dilate_compact(void* stream, const uint32_t* raw_bits, uint32_t* dil_bits, int8_t* dil_map,
                      int32_t* idx, int32_t* count, void* ws, int B, int H, int W, int kHHalf,
                      int kWHalf, int clear_raw) {
  return dilate_compact_impl(stream, raw_bits, dil_bits, dil_map, idx, count, ws, nullptr, B, H, W,
                             kHHalf, kWHalf, clear_raw);
}

int cb_dilate_compact_tiles(void* stream, const uint32_t* raw_bits, uint32_t* dil_bits,
                            int8_t* dil_map, int32_t* idx, int32_t* count, void* ws, void* tile_ws,
                            int B, int H, int W, int kHHalf, int kWHalf, int clear_raw) {
  CB_CHECK_ARG(tile_ws, "dilate_compact_tiles: null tile workspace");
  if ((long long)cb_bitmap_words(B, H, W) == 0) cudaMemsetAsync((int32_t*)tile_ws + 1, 0, sizeof(int32_t), (cudaStream_t)stream);
  return dilate_compact_impl(stream, raw_bits, dil_bits, dil_map, idx, count, ws, tile_ws, B, H, W,
                             kHHalf, kWHalf, clear_raw);
}

int cb_dilate_tiles(void* stream, const uint32_t* raw_bits, uint32_t* dil_bits, int32_t* count, void* ws,
                    void* tile_ws, int B, int H, int W, int kHHalf, int kWHalf, int clear_raw) {
  CB_CHECK_ARG(tile_ws && dil_bits, "dilate_tiles: null pointer");
  if ((long long)cb_bitmap_words(B, H, W) == 0) cudaMemsetAsync((int32_t*)tile_ws + 1, 0, sizeof(int32_t), (cudaStream_t)stream);
  // CBINFER_DILATE_TILES=1: dilate_tiles_kernel (one block per tile row, one packed atomic per block).
  // Faster alone (L1 bitmap of the bench: 4.7 vs 5.7 us back to back) but the whole step measured 1.5 %
  // SLOWER with it on two boxes (0.193 vs 0.190 ms), so dilate_compact_kernel's tile mode stays the default.
  static const bool lean = [] {
    const char* e = getenv("CBINFER_DILATE_TILES");
    return e && e[0] == '1';
  }();
  const int ty = cb::tile_grid_y(H), xp = cb::tile_grid_xp(W);
  if (lean && (long long)cb_bitmap_words(B, H, W) > 0 && cb::dilate_tiles_ok(B, H, W, kHHalf, ty, xp) &&
      kWHalf <= 31 && kHHalf >= 0 && kWHalf >= 0 && ((uintptr_t)ws % 8) == 0) {
    CB_CHECK_ARG(raw_bits && count && ws && raw_bits != dil_bits, "dilate_tiles: bad pointers");
    const int Wd = (W + 31) / 32;
    // the packed counter is the first 8 bytes of the compaction workspace (CompactHeader.reserved/done,
    // both zero at rest, so the two kernels can share one workspace)
    cb::launch_pdl(dilate_tiles_kernel, dim3((unsigned)(B * ty)), dim3(kDtThreads), cb::dilate_tiles_smem(Wd, kHHalf),
                   (cudaStream_t)stream, raw_bits, dil_bits, count, (unsigned long long*)ws, B, H, W, Wd, kHHalf,
                   kWHalf, clear_raw ? const_cast<uint32_t*>(raw_bits) : nullptr, (int32_t*)tile_ws, ty, xp);
    CB_CHECK_LAUNCH("dilate_tiles");
    return 0;
  }
  return dilate_compact_impl(stream, raw_bits, dil_bits, nullptr, nullptr, count, ws, tile_ws, B, H, W,
                             kHHalf, kWHalf, clear_raw, 1);
}

int cb_pool_compact(void* stream, const uint32_t* in_bits, uint32_t* out_bits, int32_t* idx,
                    int32_t* count, void* ws, int B, int H, int W, int oH, int oW) {
  CB_CHECK_ARG(in_bits && idx && count && ws, "pool_compact: null pointer");
  CB_CHECK_ARG(in_bits != out_bits, "pool_compact: out_bits must not alias in_bits");
  const long long nwords = (long long)cb_bitmap_words(B, oH, oW);
  cudaStream_t s = (cudaStream_t)stream;
  if (nwords == 0 || H == 0) {
    cudaMemsetAsync(count, 0, sizeof(int32_t), s);
    CB_CHECK_LAUNCH("pool_compact(memset)");
    return 0;
  }
  CB_CHECK_ARG(nwords < (1ll << 31), "pool_compact: bitmap too large");
  const int ntiles = (int)((nwords + kCompactTile - 1) / kCompactTile);
  cb::launch_pdl(dilate_compact_kernel<false>, ntiles, kCompactThreads, 0, s, in_bits, out_bits, nullptr, idx, count, ws,
                                                          B, oH, oW, (oW + 31) / 32, 0, 0,
                                                          (int)nwords, ntiles, H, (W + 31) / 32, nullptr,
                                                          nullptr, 0, 0, (int)compact_coop(), 0, cb::PrefetchHints{});
  CB_CHECK_LAUNCH("pool_compact");
  return 0;
}

int cb_change_detect_sparse(void* stream, int dtype, const void* x, long long x_sb, long long x_sc,
                            long long x_sy, long long x_sx, void* state, long long s_sb,
                            long long s_sc, long long s_sy, long long s_sx, int aux_mode,
                            void* aux_hi, void* aux_lo, const int32_t* candidates, const int32_t* n_candidates,
                            uint32_t* raw_bits, int B, int C, int H, int W, float threshold,
                            int update_mode, int bits_are_clear) {
  CB_CHECK_ARG(x && state && raw_bits && candidates && n_candidates, "change_detect_sparse: null pointer");
  CB_CHECK_ARG(B >= 0 && C > 0 && H >= 0 && W >= 0, "change_detect_sparse: bad shape");
  CB_DISPATCH_DTYPE(dtype, return (launch_detect_sparse<T, VEC>(
                               (cudaStream_t)stream, x, x_sb, x_sc, x_sy, x_sx, state, s_sb, s_sc,
                               s_sy, s_sx, aux_mode, aux_hi, aux_lo, candidates, n_candidates, raw_bits,
                               B, C, H, W,
                               threshold, update_mode, bits_are_clear)));
  return 0;
}

int cb_compact_small_max_words(void) { return cb::kCompactWin; }

int cb_change_detect_sparse_compact(void* stream, int dtype, const void* x, long long x_sb, long long x_sc,
                                    long long x_sy, long long x_sx, void* state, long long s_sb,
                                    long long s_sc, long long s_sy, long long s_sx, int aux_mode,
                                    void* aux_hi, void* aux_lo, const int32_t* candidates,
                                    const int32_t* n_candidates, uint32_t* raw_bits, int B, int C, int H,
                                    int W, float threshold, int update_mode, int bits_are_clear,
                                    uint32_t* dil_bits, int32_t* idx, int32_t* count, void* sync_ws,
                                    int kHHalf, int kWHalf, int clear_raw) {
  CB_CHECK_ARG(x && state && raw_bits && candidates && n_candidates && idx && count && sync_ws,
               "change_detect_sparse_compact: null pointer");
  CB_CHECK_ARG(B >= 0 && C > 0 && H >= 0 && W >= 0, "change_detect_sparse_compact: bad shape");
  CB_CHECK_ARG(kHHalf >= 0 && kWHalf >= 0 && kWHalf <= 31, "change_detect_sparse_compact: kWHalf must be <= 31");
  CB_CHECK_ARG(raw_bits != dil_bits, "change_detect_sparse_compact: dil_bits must not alias raw_bits");
  const long long nwords = (long long)cb_bitmap_words(B, H, W);
  if (nwords == 0) {
    cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream);
    return 0;
  }
  CB_CHECK_ARG(nwords <= cb::kCompactWin, "change_detect_sparse_compact: bitmap of %lld words > %d", nwords,
               cb::kCompactWin);
  cb::FusedCompact fc{dil_bits, idx, count, (unsigned*)sync_ws, kHHalf, kWHalf, (int)nwords, clear_raw};
  CB_DISPATCH_DTYPE(dtype, return (launch_detect_sparse<T, VEC>(
                               (cudaStream_t)stream, x, x_sb, x_sc, x_sy, x_sx, state, s_sb, s_sc,
                               s_sy, s_sx, aux_mode, aux_hi, aux_lo, candidates, n_candidates, raw_bits,
                               B, C, H, W, threshold, update_mode, bits_are_clear, &fc)));
  return 0;
}

int cb_map_to_bits(void* stream, const int8_t* map, uint32_t* bits, int B, int H, int W) {
  CB_CHECK_ARG(map && bits, "map_to_bits: null pointer");
  const long long nwords = (long long)cb_bitmap_words(B, H, W);
  if (nwords == 0) return 0;
  const long long blocks = (nwords + 7) / 8;
  cb::launch_pdl(map_to_bits_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, map, bits, H, W,
                                                                        (W + 31) / 32, nwords);
  CB_CHECK_LAUNCH("map_to_bits");
  return 0;
}

size_t cb_packed_weight_bytes(int dtype, int gemm, int Cout, int Cin, int kH, int kW) {
  const int Cp = gemm == CB_GEMM_TC_BF16X3 ? cb::pitch16_of(Cin) : cb_channel_pitch(dtype, Cin);
  if (gemm == CB_GEMM_SIMT_F32) {
    const int CoutP = (Cout + 3) / 4 * 4;
    return (size_t)kH * kW * Cp * CoutP * sizeof(float);
  }
  return cb::umma_packed_bytes(dtype, gemm, Cout, Cp, kH, kW);
}

int cb_pack_weights(void* stream, int dtype, int gemm, const void* weight, void* packed, int Cout,
                    int Cin, int kH, int kW) {
  CB_CHECK_ARG(weight && packed, "pack_weights: null pointer");
  const int Cp = gemm == CB_GEMM_TC_BF16X3 ? cb::pitch16_of(Cin) : cb_channel_pitch(dtype, Cin);
  cudaStream_t s = (cudaStream_t)stream;
  if (gemm == CB_GEMM_SIMT_F32) {
    const int CoutP = (Cout + 3) / 4 * 4;
    const long long total = (long long)kH * kW * Cp * CoutP;
    CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(pack_weights_simt_kernel<T>, grid_for(total, 256), 256, 0, s, 
                                 (const T*)weight, (float*)packed, Cout, Cin, kH, kW, Cp, CoutP)));
    CB_CHECK_LAUNCH("pack_weights");
    return 0;
  }
  return cb::umma_pack_weights(s, dtype, gemm, weight, packed, Cout, Cin, Cp, kH, kW);
}

size_t cb_conv_ws_bytes(void) {
  // flags + one fp32 partial tile per resident CTA slot; 256 KB per SM covers every tiling
  return (size_t)cb::UM_SK_FLAG_BYTES + (size_t)sm_count() * 256 * 1024;
}

static int conv_update_impl(const cb::RowMask& mk, void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                   int pitch_in, const int32_t* idx, const int32_t* count, const void* packed_w,
                   const float* bias, void* out, int pitch_out, int B, int H, int W, int Cin,
                   int Cout, int kH, int kW, int relu, void* ws, size_t ws_bytes) {
  CB_CHECK_ARG(state && idx && count && packed_w && bias && out, "conv_update: null pointer");
  const int want_pitch = gemm == CB_GEMM_TC_BF16X3 ? cb::pitch16_of(Cin) : cb_channel_pitch(dtype, Cin);
  CB_CHECK_ARG(pitch_in == want_pitch, "conv_update: pitch_in %d != channel pitch %d", pitch_in,
               want_pitch);
  CB_CHECK_ARG(pitch_out >= Cout, "conv_update: pitch_out < Cout");
  CB_CHECK_ARG((kH & 1) && (kW & 1), "conv_update: even kernel sizes unsupported (padding==k//2)");
  CB_CHECK_ARG((long long)B * H * W < (1ll << 31), "conv_update: more than 2^31 pixels");
  if (B == 0 || H == 0 || W == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  CB_CHECK_ARG(!(mk.bits && gemm == CB_GEMM_SIMT_F32), "conv_update_masked: tensor-core modes only");
  if (gemm == CB_GEMM_SIMT_F32) {
    const int CoutP = (Cout + 3) / 4 * 4;
    const unsigned grid = (unsigned)(sm_count() * 4);
    CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(conv_simt_kernel<T>, grid, SM_THREADS, 0, s, 
                                 (const T*)state, pitch_in, idx, count, (const float*)packed_w, bias,
                                 (T*)out, pitch_out, H, W, Cout, CoutP, kH, kW, relu)));
    CB_CHECK_LAUNCH("conv_update(simt)");
    return 0;
  }
  return cb::umma_conv_update(s, dtype, gemm, state, state_lo, pitch_in, idx, count, packed_w, bias, out,
                              pitch_out, B, H, W, Cin, Cout, kH, kW, relu, ws, ws_bytes, mk);
}

int cb_conv_update(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                   int pitch_in, const int32_t* idx, const int32_t* count, const void* packed_w,
                   const float* bias, void* out, int pitch_out, int B, int H, int W, int Cin,
                   int Cout, int kH, int kW, int relu, void* ws, size_t ws_bytes) {
  return conv_update_impl(cb::RowMask{nullptr, nullptr, nullptr, nullptr, 0}, stream, dtype, gemm, state,
                          state_lo, pitch_in, idx, count, packed_w, bias, out, pitch_out, B, H, W, Cin,
                          Cout, kH, kW, relu ? 1 : 0, ws, ws_bytes);
}

int cb_conv_accumulate(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                       int pitch_in, const int32_t* idx, const int32_t* count, const void* packed_w,
                       void* out, int pitch_out, int B, int H, int W, int Cin, int Cout, int kH, int kW,
                       void* ws, size_t ws_bytes) {
  CB_CHECK_ARG(gemm != CB_GEMM_SIMT_F32, "conv_accumulate: tensor-core modes only");
  // flag bit 1 of the epilogue: out += contraction, bias unused (any valid pointer satisfies the checks)
  return conv_update_impl(cb::RowMask{nullptr, nullptr, nullptr, nullptr, 0}, stream, dtype, gemm, state,
                          state_lo, pitch_in, idx, count, packed_w, (const float*)packed_w, out, pitch_out,
                          B, H, W, Cin, Cout, kH, kW, 2, ws, ws_bytes);
}

int cb_fg_detect(void* stream, const float* x, long long x_sb, long long x_sc, long long x_sy,
                 long long x_sx, float* prev, long long p_sb, long long p_sy, int p_pitch, void* delta_hi,
                 void* delta_lo, uint32_t* raw_bits, int32_t* count, int B, int C, int H, int W,
                 float threshold) {
  CB_CHECK_ARG(x && prev && delta_hi && delta_lo && raw_bits, "fg_detect: null pointer");
  CB_CHECK_ARG(B >= 0 && C > 0 && H >= 0 && W >= 0, "fg_detect: bad shape");
  return cb::launch_fg_detect((cudaStream_t)stream, x, x_sb, x_sc, x_sy, x_sx, prev, p_sb, p_sy, p_pitch,
                              delta_hi, delta_lo, raw_bits, count, B, C, H, W, threshold);
}

int cb_conv_tiled_supported(int dtype, int gemm, int B, int H, int W, int Cin, int Cout, int kH, int kW) {
  if (gemm == CB_GEMM_SIMT_F32) return 0;
  const int Cp = gemm == CB_GEMM_TC_BF16X3 ? cb::pitch16_of(Cin) : cb_channel_pitch(dtype, Cin);
  cb::TilePlan plan;
  cb::umma_tile_plan(plan, dtype, gemm, Cp, B, H, W, Cout, cb_channel_pitch(dtype, Cout), kH, kW, 0);
  if (!plan.ok) return 0;
  return (plan.mma_clk_per_tile <= cb::tile_clk_limit() || plan.short_k) ? 1 : 2;
}

static int conv_update_tiled_impl(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                         int pitch_in, const void* tile_ws, const uint32_t* dil_bits,
                         const void* packed_w, const float* bias, void* out, int pitch_out, int B,
                         int H, int W, int Cin, int Cout, int kH, int kW, int relu, const cb::PoolFuse& pf,
                         const cb::TileSelf* self = nullptr) {
  cb::TileSelf sf;
  memset(&sf, 0, sizeof(sf));
  if (self) sf = *self;
  CB_CHECK_ARG(state && tile_ws && dil_bits && packed_w && bias && out, "conv_update_tiled: null pointer");
  CB_CHECK_ARG(gemm != CB_GEMM_SIMT_F32, "conv_update_tiled: tensor-core modes only");
  const int want_pitch = gemm == CB_GEMM_TC_BF16X3 ? cb::pitch16_of(Cin) : cb_channel_pitch(dtype, Cin);
  CB_CHECK_ARG(pitch_in == want_pitch, "conv_update_tiled: pitch_in %d != channel pitch %d", pitch_in,
               want_pitch);
  CB_CHECK_ARG(pitch_out >= Cout, "conv_update_tiled: pitch_out < Cout");
  CB_CHECK_ARG(gemm != CB_GEMM_TC_BF16X3 || dtype == CB_F32, "conv_update_tiled: 3xBF16 is for fp32 data");
  if (B == 0 || H == 0 || W == 0) return 0;
  return cb::umma_conv_update_tiled((cudaStream_t)stream, dtype, gemm, state, state_lo, pitch_in,
                                    (const int32_t*)tile_ws, dil_bits, packed_w, bias, out, pitch_out,
                                    B, H, W, Cout, kH, kW, relu, pf, sf);
}

int cb_conv_update_tiled(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                         int pitch_in, const void* tile_ws, const uint32_t* dil_bits,
                         const void* packed_w, const float* bias, void* out, int pitch_out, int B,
                         int H, int W, int Cin, int Cout, int kH, int kW, int relu) {
  cb::PoolFuse pf;
  memset(&pf, 0, sizeof(pf));
  return conv_update_tiled_impl(stream, dtype, gemm, state, state_lo, pitch_in, tile_ws, dil_bits, packed_w,
                                bias, out, pitch_out, B, H, W, Cin, Cout, kH, kW, relu, pf);
}

#ifdef CB_TILE_TRACE
/* debugging aid of trace builds only (tools/tile_trace.py): copy the tile kernel's per-CTA timeline out and clear it */
int cb_debug_tile_trace(long long* dst, int n_ctas) {
  if (n_ctas > 2048) n_ctas = 2048;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(dst, cb::cb_tile_trace, (size_t)n_ctas * cb::TL_TRACE_EV * sizeof(long long));
  static long long zeros[2048 * cb::TL_TRACE_EV];
  cudaMemcpyToSymbol(cb::cb_tile_trace, zeros, sizeof(zeros));
  return cb::TL_TRACE_EV;
}
#endif

int cb_conv_tiled_pool_supported(int dtype, int gemm, int Cout) {
  (void)dtype;
  return gemm != CB_GEMM_SIMT_F32 && cb::tile_pool_ok(gemm, Cout) ? 1 : 0;
}

static int tiled_pool_impl(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                              int pitch_in, const void* tile_ws, const uint32_t* dil_bits,
                              const void* packed_w, const float* bias, void* out, int pitch_out, int B,
                              int H, int W, int Cin, int Cout, int kH, int kW, int relu,
                              void* pool_out, long long o_sb, long long o_sy, int o_pitch, int oH, int oW,
                              void* next_state, long long n_sb, long long n_sy, int n_pitch, int aux_mode,
                              void* aux_hi, void* aux_lo, uint32_t* next_raw_bits, float threshold,
                              int update_mode, const cb::TileSelf* self) {
  CB_CHECK_ARG(pool_out && next_state && next_raw_bits, "conv_update_tiled_pool: null pointer");
  CB_CHECK_ARG(update_mode == CB_UPDATE_NONE || update_mode == CB_UPDATE_CHANGED || update_mode == CB_UPDATE_ALL,
               "conv_update_tiled_pool: bad update_mode %d", update_mode);
  const size_t es = cb::esize(dtype);
  const int vec = (int)(16 / es);
  const bool ok = (o_pitch % vec) == 0 && o_pitch == n_pitch && o_pitch >= Cout && pitch_out == o_pitch &&
                  ((o_sy * es) % 16) == 0 && ((n_sy * es) % 16) == 0 && ((o_sb * es) % 16) == 0 &&
                  ((n_sb * es) % 16) == 0 && ((uintptr_t)pool_out % 16) == 0 && ((uintptr_t)next_state % 16) == 0 &&
                  ((uintptr_t)out % 16) == 0;
  CB_CHECK_ARG(ok, "conv_update_tiled_pool: needs pixel-major, 16-byte aligned maps of equal pitch");
  cb::PoolFuse pf;
  memset(&pf, 0, sizeof(pf));
  pf.out = pool_out; pf.o_sb = o_sb; pf.o_sy = o_sy; pf.op = o_pitch; pf.oH = oH; pf.oW = oW;
  pf.nst = next_state; pf.n_sb = n_sb; pf.n_sy = n_sy; pf.np = n_pitch;
  pf.nbits = next_raw_bits; pf.thr = threshold; pf.update = update_mode;
  int rc = 0;
  CB_DISPATCH_DTYPE(dtype, rc = cb::make_aux<T>(pf.aux, aux_mode, aux_hi, aux_lo, next_state, Cout));
  if (rc) return rc;
  return conv_update_tiled_impl(stream, dtype, gemm, state, state_lo, pitch_in, tile_ws, dil_bits, packed_w,
                                bias, out, pitch_out, B, H, W, Cin, Cout, kH, kW, relu, pf, self);
}

int cb_conv_update_tiled_pool(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                              int pitch_in, const void* tile_ws, const uint32_t* dil_bits,
                              const void* packed_w, const float* bias, void* out, int pitch_out, int B,
                              int H, int W, int Cin, int Cout, int kH, int kW, int relu,
                              void* pool_out, long long o_sb, long long o_sy, int o_pitch, int oH, int oW,
                              void* next_state, long long n_sb, long long n_sy, int n_pitch, int aux_mode,
                              void* aux_hi, void* aux_lo, uint32_t* next_raw_bits, float threshold,
                              int update_mode) {
  return tiled_pool_impl(stream, dtype, gemm, state, state_lo, pitch_in, tile_ws, dil_bits, packed_w, bias, out,
                         pitch_out, B, H, W, Cin, Cout, kH, kW, relu, pool_out, o_sb, o_sy, o_pitch, oH, oW,
                         next_state, n_sb, n_sy, n_pitch, aux_mode, aux_hi, aux_lo, next_raw_bits, threshold,
                         update_mode, nullptr);
}

int cb_conv_tiled_self_supported(int B, int H, int W, int kH, int kW) {
  // one warp lane per raw row of a 16-row tile's window (16 + 2*kHHalf <= 32), funnel shifts up to 31;
  // the barrier word carries the tile count in 20 bits
  return kH >= 1 && kW >= 1 && (kH & 1) && (kW & 1) && (kH - 1) / 2 <= 8 && (kW - 1) / 2 <= 31 &&
                 B >= 0 && H >= 0 && W >= 0 && (long long)B * H * W < (1ll << 31) &&
                 (long long)B * cb::tile_grid_y(H) * cb::tile_grid_xp(W) < (1ll << 20)
             ? 1 : 0;
}

int cb_conv_update_tiled_self(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                              int pitch_in, void* tile_ws, uint32_t* dil_bits,
                              const void* packed_w, const float* bias, void* out, int pitch_out, int B,
                              int H, int W, int Cin, int Cout, int kH, int kW, int relu,
                              void* pool_out, long long o_sb, long long o_sy, int o_pitch, int oH, int oW,
                              void* next_state, long long n_sb, long long n_sy, int n_pitch, int aux_mode,
                              void* aux_hi, void* aux_lo, uint32_t* next_raw_bits, float threshold,
                              int update_mode, const uint32_t* raw_bits, int32_t* count, void* ws,
                              int clear_raw) {
  CB_CHECK_ARG(raw_bits && count && ws && dil_bits && tile_ws, "conv_update_tiled_self: null pointer");
  CB_CHECK_ARG(raw_bits != dil_bits, "conv_update_tiled_self: dil_bits must not alias raw_bits");
  CB_CHECK_ARG(cb_conv_tiled_self_supported(B, H, W, kH, kW),
               "conv_update_tiled_self: filter %dx%d / map %dx%dx%d not supported", kH, kW, B, H, W);
  if (B == 0 || H == 0 || W == 0) {
    cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream);
    cudaMemsetAsync((int32_t*)tile_ws + 1, 0, sizeof(int32_t), (cudaStream_t)stream);
    CB_CHECK_LAUNCH("conv_update_tiled_self(memset)");
    return 0;
  }
  cb::TileSelf sf;
  sf.raw = raw_bits;
  sf.dil = dil_bits;
  sf.clear = clear_raw ? const_cast<uint32_t*>(raw_bits) : nullptr;
  sf.count = count;
  sf.acc_pix = &reinterpret_cast<cb::CompactHeader*>(ws)->reserved;
  if (!pool_out) {
    cb::PoolFuse pf;
    memset(&pf, 0, sizeof(pf));
    return conv_update_tiled_impl(stream, dtype, gemm, state, state_lo, pitch_in, tile_ws, dil_bits, packed_w,
                                  bias, out, pitch_out, B, H, W, Cin, Cout, kH, kW, relu, pf, &sf);
  }
  return tiled_pool_impl(stream, dtype, gemm, state, state_lo, pitch_in, tile_ws, dil_bits, packed_w, bias, out,
                         pitch_out, B, H, W, Cin, Cout, kH, kW, relu, pool_out, o_sb, o_sy, o_pitch, oH, oW,
                         next_state, n_sb, n_sy, n_pitch, aux_mode, aux_hi, aux_lo, next_raw_bits, threshold,
                         update_mode, &sf);
}

int cb_conv_update_masked(void* stream, int dtype, int gemm, const void* state, const void* state_lo,
                          int pitch_in, const int32_t* idx, const int32_t* count, const void* packed_w,
                          const float* bias, void* out, int pitch_out, int B, int H, int W, int Cin,
                          int Cout, int kH, int kW, int relu, void* ws, size_t ws_bytes,
                          uint32_t* mask_bits, int clear_mask, int32_t* count_out, void* sync_ws) {
  CB_CHECK_ARG(mask_bits && sync_ws, "conv_update_masked: null pointer");
  cb::RowMask mk;
  mk.bits = mask_bits;
  mk.clear = clear_mask ? mask_bits : nullptr;
  mk.count_out = count_out;
  mk.sync = (unsigned*)sync_ws;
  mk.nwords = (int)cb_bitmap_words(B, H, W);
  return conv_update_impl(mk, stream, dtype, gemm, state, state_lo, pitch_in, idx, count, packed_w, bias,
                          out, pitch_out, B, H, W, Cin, Cout, kH, kW, relu ? 1 : 0, ws, ws_bytes);
}

int cb_tail_supported(int dtype, int gemm, int C0, int C1, int C2) {
  return dtype == CB_F32 && gemm == CB_GEMM_TC_BF16X3 && cb::tail_supported(C0, C1, C2) ? 1 : 0;
}

int cb_tail_update(void* stream, const float* x, float* state1, const void* packed_w1, const float* bias1,
                   float* out1, int relu1, float thr1, float* state2, const void* packed_w2,
                   const float* bias2, float* out2, int pitch_out2, int relu2, float thr2,
                   const int32_t* candidates, const int32_t* n_candidates, int C0, int C1, int C2,
                   int update_mode, int32_t* count1, int32_t* count2, void* sync_ws) {
  CB_CHECK_ARG(x && state1 && packed_w1 && bias1 && out1 && state2 && packed_w2 && bias2 && out2 && candidates &&
                   n_candidates && count1 && count2 && sync_ws, "tail_update: null pointer");
  CB_CHECK_ARG(update_mode == CB_UPDATE_NONE || update_mode == CB_UPDATE_CHANGED || update_mode == CB_UPDATE_ALL,
               "tail_update: bad update_mode %d", update_mode);
  CB_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)state1 % 16) == 0 && ((uintptr_t)out1 % 16) == 0 &&
                   ((uintptr_t)state2 % 16) == 0 && ((uintptr_t)out2 % 16) == 0 &&
                   ((uintptr_t)packed_w1 % 128) == 0 && ((uintptr_t)packed_w2 % 128) == 0,
               "tail_update: maps must be 16-byte, packed weights 128-byte aligned");
  cb::TailArgs a;
  a.x = x; a.st1 = state1; a.out1 = out1; a.st2 = state2; a.out2 = out2;
  a.cand = candidates; a.ncand = n_candidates; a.bias1 = bias1; a.bias2 = bias2;
  a.count1 = count1; a.count2 = count2; a.sync = (unsigned*)sync_ws;
  a.xp = C0; a.p1 = C1; a.p2 = pitch_out2; a.C0 = C0; a.C1 = C1; a.C2 = C2;
  a.relu1 = relu1; a.relu2 = relu2; a.update = update_mode; a.thr1 = thr1; a.thr2 = thr2;
  static const int tail_adapt = [] {
    const char* e = getenv("CBINFER_TAIL_ADAPT");            // tuning knob: 0 = always 128 rows per tile
    return (e && e[0] == '0') ? 0 : 1;
  }();
  a.adapt = tail_adapt;
  return cb::tail_update((cudaStream_t)stream, a, packed_w1, packed_w2);
}

int cb_maxpool2x2(void* stream, int dtype, const void* x, long long x_sb, long long x_sc,
                  long long x_sy, long long x_sx, const int32_t* idx, const int32_t* count,
                  const uint32_t* dil_bits, void* out, long long o_sb, long long o_sc,
                  long long o_sy, long long o_sx, int B, int C, int H, int W, int oH, int oW) {
  CB_CHECK_ARG(x && idx && count && out, "maxpool2x2: null pointer");
  if (B == 0 || H == 0 || W == 0 || C == 0) return 0;
  const unsigned grid = (unsigned)(sm_count() * 16);
  const size_t es = cb::esize(dtype);
  const int vec = (int)(16 / es);
  // pixel-major fast path: channel-contiguous, 16-byte aligned pixels on both sides.  Chunks that
  // cover pad channels (pitch > C) are processed too: pads are zero on both sides by construction.
  const bool vec_ok = x_sc == 1 && o_sc == 1 && (x_sx % vec) == 0 && (o_sx % vec) == 0 &&
                      x_sx == o_sx && x_sx >= C && x_sx < (1 << 20) && ((x_sy * es) % 16) == 0 &&
                      ((o_sy * es) % 16) == 0 && ((x_sb * es) % 16) == 0 && ((o_sb * es) % 16) == 0 &&
                      ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
                      (x_sx == C || (C + vec - 1) / vec * vec == x_sx);
  if (vec_ok) {
    const int cpp = (int)(x_sx / vec);
    int glog = 0;
    while ((1 << glog) < cpp && glog < 5) ++glog;
    glog = cb::glog_tuned(glog);
    CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(maxpool2x2_vec_kernel<T, VEC>, grid, 256, 0, (cudaStream_t)stream, 
                                 (const T*)x, x_sb, x_sy, (int)x_sx, idx, count, dil_bits, (T*)out,
                                 o_sb, o_sy, (int)o_sx, cpp, glog, H, W, oH, oW)));
  } else {
    CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(maxpool2x2_kernel<T>, grid, 256, 0, (cudaStream_t)stream, 
                                 (const T*)x, x_sb, x_sc, x_sy, x_sx, idx, count, dil_bits, (T*)out,
                                 o_sb, o_sc, o_sy, o_sx, C, H, W, oH, oW)));
  }
  CB_CHECK_LAUNCH("maxpool2x2");
  return 0;
}

size_t cb_detect_compact_ws_bytes(int B, int H, int W) { return 16 + (size_t)B * H * W + 16; }

int cb_detect_compact_sparse(void* stream, int dtype, const void* x, long long x_sb, long long x_sc,
                             long long x_sy, long long x_sx, void* state, long long s_sb,
                             long long s_sc, long long s_sy, long long s_sx, int aux_mode,
                             void* aux_hi, void* aux_lo, const int32_t* candidates,
                             const int32_t* n_candidates, int32_t* idx, int32_t* count,
                             uint32_t* bits, void* ws, int B, int C, int H, int W, float threshold,
                             int update_mode) {
  CB_CHECK_ARG(x && state && candidates && n_candidates && idx && count && ws,
               "detect_compact_sparse: null pointer");
  CB_CHECK_ARG(idx != candidates, "detect_compact_sparse: idx must not alias candidates");
  CB_DISPATCH_DTYPE(dtype, return (launch_detect_compact_sparse<T, VEC>(
                               (cudaStream_t)stream, x, x_sb, x_sc, x_sy, x_sx, state, s_sb, s_sc,
                               s_sy, s_sx, aux_mode, aux_hi, aux_lo, candidates, n_candidates, idx,
                               count, bits, ws, B, C, H, W, threshold, update_mode)));
  return 0;
}

int cb_maxpool2x2_detect(void* stream, int dtype, const void* x, long long x_sb, long long x_sy,
                         int x_pitch, const int32_t* idx, const int32_t* count,
                         const uint32_t* dil_bits, void* out, long long o_sb, long long o_sy,
                         int o_pitch, int B, int C, int H, int W, int oH, int oW, void* next_state,
                         long long n_sb, long long n_sy, int n_pitch, int aux_mode, void* aux_hi,
                         void* aux_lo, uint32_t* next_raw_bits, float threshold, int update_mode) {
  CB_CHECK_ARG(x && idx && count && dil_bits && out && next_state && next_raw_bits,
               "maxpool2x2_detect: null pointer");
  if (B == 0 || H == 0 || W == 0 || C == 0) return 0;
  const size_t es = cb::esize(dtype);
  const int vec = (int)(16 / es);
  const bool ok = (x_pitch % vec) == 0 && x_pitch == o_pitch && o_pitch == n_pitch && x_pitch >= C &&
                  (C + vec - 1) / vec * vec == x_pitch && ((x_sy * es) % 16) == 0 &&
                  ((o_sy * es) % 16) == 0 && ((n_sy * es) % 16) == 0 && ((x_sb * es) % 16) == 0 &&
                  ((o_sb * es) % 16) == 0 && ((n_sb * es) % 16) == 0 && ((uintptr_t)x % 16) == 0 &&
                  ((uintptr_t)out % 16) == 0 && ((uintptr_t)next_state % 16) == 0;
  CB_CHECK_ARG(ok, "maxpool2x2_detect: needs pixel-major, 16-byte aligned tensors of equal pitch");
  const int cpp = x_pitch / vec;
  int glog = 0;
  while ((1 << glog) < cpp && glog < 5) ++glog;
  glog = cb::glog_tuned(glog);
  const unsigned grid = (unsigned)(sm_count() * 16);
  cudaStream_t s = (cudaStream_t)stream;
#define CB_MPD(U_)                                                                                \
  CB_DISPATCH_DTYPE(dtype, {                                                                      \
    AuxPlanes aux;                                                                                \
    if (int rc = make_aux<T>(aux, aux_mode, aux_hi, aux_lo, next_state, C)) return rc;            \
    cb::launch_pdl(maxpool2x2_detect_kernel<T, VEC, U_>, grid, 256, 0, s, (const T*)x, x_sb,     \
                   x_sy, x_pitch, idx, count, dil_bits, (T*)out, o_sb, o_sy, o_pitch, cpp, glog, \
                   H, W, oH, oW, (T*)next_state, n_sb, n_sy, n_pitch, aux, next_raw_bits,         \
                   thr_cast<T>(threshold));                                                       \
  })
  switch (update_mode) {
    case CB_UPDATE_NONE: CB_MPD(CB_UPDATE_NONE); break;
    case CB_UPDATE_CHANGED: CB_MPD(CB_UPDATE_CHANGED); break;
    case CB_UPDATE_ALL: CB_MPD(CB_UPDATE_ALL); break;
    default: return cb::fail(2, "maxpool2x2_detect: bad update_mode %d", update_mode);
  }
#undef CB_MPD
  CB_CHECK_LAUNCH("maxpool2x2_detect");
  return 0;
}

int cb_gen_xmatrix(void* stream, int dtype, void* columns, const void* input, const int32_t* idx,
                   int kW, int kH, int C, int W, int H, int n) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(columns && input && idx, "gen_xmatrix: null pointer");
  const long long total = (long long)n * C * kH * kW;
  CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(gen_xmatrix_kernel<T>, grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream, 
                               (T*)columns, (const T*)input, idx, kW, kH, C, W, H, n)));
  CB_CHECK_LAUNCH("gen_xmatrix");
  return 0;
}

int cb_matrix_mult(void* stream, int dtype, const void* X, const void* weight, const void* bias,
                   void* Y, int n, int K, int Cout) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(X && weight && bias && Y, "matrix_mult: null pointer");
  dim3 grid((Cout + 31) / 32, (n + 31) / 32);
  CB_CHECK_ARG(grid.y < 65536, "matrix_mult: n too large for the staged op");
  CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(matrix_mult_kernel<T>, grid, 256, 0, (cudaStream_t)stream, 
                               (const T*)X, (const T*)weight, (const T*)bias, (T*)Y, n, K, Cout)));
  CB_CHECK_LAUNCH("matrix_mult");
  return 0;
}

int cb_update_output(void* stream, int dtype, const void* Yt, void* output, const int32_t* idx,
                     int numOutputPixel, int n, int Cout, int relu) {
  if (n <= 0) return 0;
  CB_CHECK_ARG(Yt && output && idx, "update_output: null pointer");
  const long long total = (long long)n * Cout;
  CB_DISPATCH_DTYPE(dtype, (cb::launch_pdl(update_output_kernel<T>, grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream, 
                               (const T*)Yt, (T*)output, idx, numOutputPixel, n, Cout, relu)));
  CB_CHECK_LAUNCH("update_output");
  return 0;
}

int cb_fg_update(void* stream, const float* x, float* prev, const float* weight, float* out,
                 int32_t* count, int B, int Cin, int Cout, int H, int W, int kH, int kW,
                 float threshold) {
  CB_CHECK_ARG(x && prev && weight && out && count, "fg_update: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(count, 0, sizeof(int32_t), s);
  const long long total = (long long)B * Cin * H * W;
  if (total == 0) return 0;
  cb::launch_pdl(fg_update_kernel, grid_for((total + 31) / 32 * 32, 256, 8), 256, 0, s, x, prev, weight, out, count, B, Cin,
                                                             Cout, H, W, kH, kW, threshold);
  CB_CHECK_LAUNCH("fg_update");
  return 0;
}

/* ---- frame ingest: resizing on the device (ingest.cuh) -------------------------------------------- */
size_t cb_resize_ws_bytes(int sH, int sW, int dH, int dW, int C) {
  if (sH <= 0 || sW <= 0 || dH <= 0 || dW <= 0 || C <= 0) return 0;
  return cb::resize_plan(sH, sW, dH, dW, C).bytes;
}

int cb_resize_bicubic_u8_init(void* stream, void* ws, int sH, int sW, int dH, int dW, int C) {
  CB_CHECK_ARG(ws && sH > 0 && sW > 0 && dH > 0 && dW > 0 && C >= 1 && C <= 4,
               "resize_bicubic_u8_init: bad arguments (1..4 channels)");
  const cb::ResizePlan p = cb::resize_plan(sH, sW, dH, dW, C);
  CB_CHECK_ARG(p.off_tmp < (1ull << 31), "resize_bicubic_u8_init: image too large");
  // the weight tables are a function of the four sizes only: computed here in double exactly as Pillow's
  // precompute_coeffs / normalize_coeffs_8bpc do, copied once; cb_resize_bicubic_u8 is then launch-only
  std::vector<int32_t> host(p.off_tmp / 4, 0);
  host[0] = p.ksize_x; host[1] = p.ksize_y; host[2] = sH; host[3] = sW; host[4] = dH; host[5] = dW; host[6] = C;
  cb::resize_coeffs(sW, dW, p.ksize_x, host.data() + p.off_bx / 4, host.data() + p.off_kx / 4);
  cb::resize_coeffs(sH, dH, p.ksize_y, host.data() + p.off_by / 4, host.data() + p.off_ky / 4);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemcpyAsync(ws, host.data(), p.off_tmp / 4 * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return cb::fail(3, "resize_bicubic_u8_init: copying the weight tables failed: %s",
                    cudaGetErrorString(cudaGetLastError()));
  return 0;
}

int cb_resize_bicubic_u8(void* stream, const uint8_t* src, long long s_y, long long s_x, long long s_c,
                         uint8_t* dst, long long d_y, long long d_x, long long d_c, void* ws, int sH, int sW,
                         int dH, int dW, int C) {
  CB_CHECK_ARG(src && dst && ws && sH > 0 && sW > 0 && dH > 0 && dW > 0 && C >= 1 && C <= 4,
               "resize_bicubic_u8: bad arguments (1..4 channels)");
  const cb::ResizePlan p = cb::resize_plan(sH, sW, dH, dW, C);
  char* w = (char*)ws;
  uint8_t* tmp = (uint8_t*)(w + p.off_tmp);                    // [sH, dW, C] interleaved
  cudaStream_t s = (cudaStream_t)stream;
  cb::launch_pdl(cb::resize_pass_u8_kernel<1>, dim3((unsigned)((dW + 255) / 256), (unsigned)sH), dim3(256), 0, s, src,
                 s_y, s_x, s_c, tmp, (long long)dW * C, (long long)C, 1ll, sH, dW, C,
                 (const int32_t*)(w + p.off_bx), (const int32_t*)(w + p.off_kx), p.ksize_x);
  CB_CHECK_LAUNCH("resize_bicubic_u8(horizontal)");
  cb::launch_pdl(cb::resize_pass_u8_kernel<0>, dim3((unsigned)((dW + 255) / 256), (unsigned)dH), dim3(256), 0, s,
                 (const uint8_t*)tmp, (long long)dW * C, (long long)C, 1ll, dst, d_y, d_x, d_c, dH, dW, C,
                 (const int32_t*)(w + p.off_by), (const int32_t*)(w + p.off_ky), p.ksize_y);
  CB_CHECK_LAUNCH("resize_bicubic_u8(vertical)");
  return 0;
}

int cb_resize_bilinear_u8(void* stream, const uint8_t* src, long long s_y, long long s_x, long long s_c, int sH,
                          int sW, float* dst, long long d_y, long long d_x, long long d_c, int dH, int dW, int C,
                          float divisor, float cval, float clip_lo, float clip_hi) {
  CB_CHECK_ARG(src && dst && sH > 0 && sW > 0 && dH > 0 && dW > 0 && C >= 1 && divisor != 0.f,
               "resize_bilinear_u8: bad arguments");
  cb::launch_pdl(cb::resize_bilinear_kernel, dim3((unsigned)((dW + 255) / 256), (unsigned)dH), dim3(256), 0,
                 (cudaStream_t)stream, src, s_y, s_x, s_c, sH, sW, dst, d_y, d_x, d_c, dH, dW, C, (double)divisor,
                 (double)cval, (double)clip_lo, (double)clip_hi);
  CB_CHECK_LAUNCH("resize_bilinear_u8");
  return 0;
}

}  // extern "C"
