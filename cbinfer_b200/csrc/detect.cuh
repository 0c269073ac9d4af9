// detect.cuh -- change detection: thresholded |x_t - x_{t-1}| -> 1 bit per pixel.
//
// Replaces the detection half of changeDetection[_1x1]_kernel
// (reference pycbinfer/cbconv2d_cg_backend.cu:6-81, half: cbconv2d_cg_half_backend.cu:10-88).
//
// One bitmap word = 32 consecutive pixels of one image row.
//   * vector path (pixel-major x and state): WPW warps share one word; each streams its slice of
//     the 32 pixels' channels as 16-byte chunks, fully coalesced, with 2*U loads in flight per
//     lane; per-chunk flags are folded to per-pixel flags with __ballot_sync + bit-range masks and
//     OR-ed across the word's warps through shared memory.
//   * narrow path (any x strides, state pixel = one 16-byte chunk, e.g. the RGB input frame).
//   * generic path (any strides on both sides).
// HBM-bound: algorithmic bytes = 2*C*P*s read (+ P/8 bitmap, + feedback writes).
#pragma once
#include "cb_common.cuh"
#include "compact.cuh"

namespace cb {

template <typename T, int VEC>
__device__ __forceinline__ uint4 merge_tail(uint4 xv, const uint4& sv, int tail) {
  T* xe = reinterpret_cast<T*>(&xv);
  const T* se = reinterpret_cast<const T*>(&sv);
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    if (e >= tail) xe[e] = se[e];
  return xv;
}

// Auxiliary operand planes kept in step with the state (written wherever the state is written), so
// the contraction's operand split is paid once per accepted pixel instead of once per gathered tap:
//   mode 1: fp32 plane  lo = v - trunc_tf32(v)            (3xTF32: the raw state is the "hi" operand,
//           the tensor core ignores the 13 low mantissa bits)
//   mode 2: two bf16 planes  hi = bf16(v), lo = bf16(v - hi)   (3xBF16: ~16 mantissa bits at twice
//           the tensor rate and half the bytes of 3xTF32), pixel-major with their own pitch.
struct AuxPlanes {
  int mode;
  int pitch16;               // mode 2: bf16 elements per pixel (pitch16_of(C))
  long long lo_off;          // mode 1: element offset state -> lo plane (same strides as the state)
  __nv_bfloat16* hi16;       // mode 2
  __nv_bfloat16* lo16;
};

__device__ __forceinline__ unsigned tf32_lo(unsigned v) {
  return __float_as_uint(__uint_as_float(v) - __uint_as_float(v & 0xFFFFE000u));
}
__device__ __forceinline__ void bf16_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// pix = global pixel index, ch0 = first channel of this 16-byte chunk
template <typename T>
__device__ __forceinline__ void store_state(T* sptr, const uint4& v, const AuxPlanes& aux,
                                            long long pix, int ch0) {
  st16(sptr, v);
  if (sizeof(T) == 4 && aux.mode == 1) {
    st16(sptr + aux.lo_off, make_uint4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)));
  } else if (sizeof(T) == 4 && aux.mode == 2) {
    __nv_bfloat16 h[4], l[4];
    bf16_split(__uint_as_float(v.x), h[0], l[0]);
    bf16_split(__uint_as_float(v.y), h[1], l[1]);
    bf16_split(__uint_as_float(v.z), h[2], l[2]);
    bf16_split(__uint_as_float(v.w), h[3], l[3]);
    const long long o = pix * aux.pitch16 + ch0;
    *reinterpret_cast<uint2*>(aux.hi16 + o) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(aux.lo16 + o) = *reinterpret_cast<uint2*>(l);
  }
}
__device__ __forceinline__ void store_state_scalar(float* sptr, float v, const AuxPlanes& aux,
                                                   long long pix, int ch) {
  *sptr = v;
  if (aux.mode == 1) {
    sptr[aux.lo_off] = __uint_as_float(tf32_lo(__float_as_uint(v)));
  } else if (aux.mode == 2) {
    __nv_bfloat16 h, l;
    bf16_split(v, h, l);
    aux.hi16[pix * aux.pitch16 + ch] = h;
    aux.lo16[pix * aux.pitch16 + ch] = l;
  }
}
template <typename T>
__device__ __forceinline__ void store_state_scalar(T* sptr, T v, const AuxPlanes&, long long, int) {
  *sptr = v;
}

// streaming (evict-first) loads for data a frame scan reads exactly once per step: the input frame and
// the first layer's state (70 MB per step at 8 streams) should not push the small control structures
// and the maps the following kernels reuse out of L2
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ __half ld_stream(const __half* p) {
  const unsigned short v = __ldcs(reinterpret_cast<const unsigned short*>(p));
  return __ushort_as_half(v);
}
__device__ __forceinline__ __nv_bfloat16 ld_stream(const __nv_bfloat16* p) {
  const unsigned short v = __ldcs(reinterpret_cast<const unsigned short*>(p));
  return __ushort_as_bfloat16(v);
}
__device__ __forceinline__ uint4 ld16_stream(const void* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }

constexpr int kDetWarps = 8;                   // warps per block

// x: pixel-major, pitch xp (elements); state: pixel-major, pitch sp.  Rows may be strided (sy).
// U chunk-iterations are batched: all 2*U 16-byte loads of a batch are issued before the first
// ballot, so every lane keeps 2*U loads in flight (the ballots would otherwise serialise them).
// wlog = log2(warps per word): a block of 8 warps covers 8 >> wlog words.
template <typename T, int VEC, int UPDATE, int U>
__global__ void __launch_bounds__(kDetWarps * 32)
detect_vec_kernel(const T* __restrict__ x, long long x_sb, long long x_sy, int xp,
                  T* __restrict__ st, long long s_sb, long long s_sy, int sp, AuxPlanes aux,
                  uint32_t* __restrict__ bits, int B, int H, int W, int C, int Wd, T thr,
                  unsigned cpv_magic, int wlog) {
  pdl_prologue();
  __shared__ unsigned s_word[kDetWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wpw = 1 << wlog;
  const int part = wid & (wpw - 1), slot = wid >> wlog;
  const int word = blockIdx.x * (kDetWarps >> wlog) + slot;      // < 2^31 words (host-checked)
  const bool active = word < B * H * Wd;
  if (wlog > 0) {
    if (threadIdx.x < kDetWarps) s_word[threadIdx.x] = 0u;
    __syncthreads();
  } else if (!active) {
    return;
  }
  const int cpv = (C + VEC - 1) / VEC;       // 16-byte chunks per pixel that hold real channels
  const int tail = C % VEC;                  // valid elements of the last chunk (0 = all)
  // q / cpv without a divide: exact for q*cpv < 2^32 (q <= 32*cpv here)
  auto pixel_of = [&](int q) { return cpv == 1 ? q : (int)__umulhi((unsigned)q, cpv_magic); };
  const T* xb = nullptr;
  T* sb = nullptr;
  int nq = 0;
  long long pixbase = 0;
  unsigned wordbits = 0;
  if (active) {
    const int j = word % Wd;
    const int r = word / Wd;
    const int y = r % H;
    const int b = r / H;
    const int x0 = j * 32;
    const int npx = min(32, W - x0);
    xb = x + b * x_sb + y * x_sy + (long long)x0 * xp;
    sb = st + b * s_sb + y * s_sy + (long long)x0 * sp;
    nq = npx * cpv;
    pixbase = ((long long)b * H + y) * W + x0;

    bool mychg = false;                      // flag of pixel `lane` (this warp's slice only)
    const int plo = lane * cpv, phi = plo + cpv;
    for (int q0 = part * 32 * U; q0 < nq; q0 += wpw * 32 * U) {
      uint4 xv[U], sv[U];
      T* sptr[U];
      int pxs[U], ccs[U];
      bool valid[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = q0 + u * 32 + lane;
        valid[u] = q < nq;
        const int px = pixel_of(q), cc = q - px * cpv;
        pxs[u] = px;
        ccs[u] = cc;
        sptr[u] = sb + (long long)px * sp + cc * VEC;
        if (valid[u]) {
          xv[u] = ldg16(xb + (long long)px * xp + cc * VEC);
          sv[u] = ld16(sptr[u]);
          if (tail && cc == cpv - 1) xv[u] = merge_tail<T, VEC>(xv[u], sv[u], tail);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bool f = false;
        if (valid[u]) {
          f = Chunk<T>::changed(sv[u], xv[u], thr);
          if (UPDATE == CB_UPDATE_ALL) store_state<T>(sptr[u], xv[u], aux, pixbase + pxs[u], ccs[u] * VEC);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        const int qb = q0 + u * 32;
        const int lo = max(plo, qb) - qb, hi = min(phi, qb + 32) - qb;
        if (hi > lo) {
          const unsigned m = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
          mychg |= (bal & m) != 0u;
        }
      }
    }
    wordbits = __ballot_sync(0xffffffffu, mychg);
  }
  if (wlog > 0) {                              // OR the partial words of the warps sharing a word
    if (active && lane == 0 && wordbits) atomicOr(&s_word[slot], wordbits);
    __syncthreads();
    wordbits = s_word[slot];
    if (!active) return;
  }
  if (part == 0 && lane == 0) bits[word] = wordbits;

  if (UPDATE == CB_UPDATE_CHANGED && wordbits) {  // feedback: accept the new value at changed pixels
    for (int q = part * 32 + lane; q < nq; q += wpw * 32) {
      const int px = pixel_of(q), cc = q - px * cpv;
      if ((wordbits >> px) & 1u) {
        uint4 xv = ldg16(xb + (long long)px * xp + cc * VEC);
        T* sp2 = sb + (long long)px * sp + cc * VEC;
        if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, ld16(sp2), tail);
        store_state<T>(sp2, xv, aux, pixbase + px, cc * VEC);
      }
    }
  }
}

// Generic x (any strides, e.g. the user's planar NCHW frame) against a pixel-major state whose
// pixel is exactly one 16-byte chunk (C <= 4 fp32 / C <= 8 half): lane = pixel, the state is one
// vector load/store per pixel and x is read plane by plane (coalesced across lanes).
template <typename T, int VEC, int UPDATE>
__global__ void __launch_bounds__(256)
detect_narrow_kernel(const T* __restrict__ x, long long x_sb, long long x_sc, long long x_sy,
                     long long x_sx, T* __restrict__ st, long long s_sb, long long s_sy,
                     AuxPlanes aux, uint32_t* __restrict__ bits, int B, int H, int W, int C,
                     int Wd, T thr) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  // 32-bit index math (the host checks words < 2^31): two 64-bit divisions per warp were most of this
  // kernel's instructions (ncu: 173 warp instructions per 32 pixels, SM 54 % busy on an HBM-bound scan)
  const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= (unsigned)(B * H * Wd)) return;
  const unsigned r = warp / (unsigned)Wd;
  const int j = (int)(warp - r * (unsigned)Wd);
  const int b = (int)(r / (unsigned)H);
  const int y = (int)(r - (unsigned)b * (unsigned)H);
  const int xx = j * 32 + lane;
  bool f = false;
  if (xx < W) {
    const T* xp = x + b * x_sb + y * x_sy + xx * x_sx;
    T* sp = st + b * s_sb + y * s_sy + (long long)xx * VEC;
    uint4 sv = ld16_stream(sp);
    uint4 nv = sv;                             // pad lanes keep the state's (zero) value
    T* ne = reinterpret_cast<T*>(&nv);
#pragma unroll
    for (int c = 0; c < VEC; ++c)
      if (c < C) ne[c] = ld_stream(xp + c * x_sc);
    f = Chunk<T>::changed(sv, nv, thr);
    if (UPDATE == CB_UPDATE_ALL || (UPDATE == CB_UPDATE_CHANGED && f))
      store_state<T>(sp, nv, aux, ((long long)b * H + y) * W + xx, 0);
  }
  const unsigned word = __ballot_sync(0xffffffffu, f);
  if (lane == 0) bits[warp] = word;
}

// Planar x (NCHW: the pixels of a row are contiguous, any channel stride -- what a dense layer or the
// user hands in) against a pixel-major state of more than one 16-byte chunk per pixel.  The two
// layouts want opposite lane mappings (x coalesces across pixels, the state across channels), so a
// block of 8 warps transposes `nw` bitmap words (32 pixels each) through shared memory:
//   phase 1: lane = pixel, warp = (word, channel chunk): VEC coalesced row loads per chunk -> one
//            16-byte shared-memory store at [pixel][chunk] (odd chunk pitch: conflict-free);
//   phase 2: thread = (pixel, chunk) in state order: 16-byte state load (coalesced), compare, state
//            / operand-plane store, pixel flags OR-ed into the word.
// nw is chosen so that every warp has ~4 (word, chunk) pairs: 4*VEC row loads in flight per lane.
// Both sides stream at full sector efficiency; the generic kernel reads the state at a stride of one
// pixel per lane (measured 1 TB/s at C = 64).
constexpr int kPlanarMaxWords = 8;
struct PlanarWord {
  long long xoff, soff, pix;     // element offsets of the word's first pixel in x / the state, global pixel index
  int npx;                       // valid pixels (0: word beyond the map)
};

__device__ __forceinline__ void planar_words_setup(PlanarWord* wi, unsigned* s_word, int word0, int nw,
                                                   long long nwords, int H, int W, int Wd, long long x_sb,
                                                   long long x_sy, long long x_sx, long long s_sb,
                                                   long long s_sy, int sp) {
  if (threadIdx.x < (unsigned)nw) {
    const long long word = (long long)word0 + threadIdx.x;
    PlanarWord w = {0, 0, 0, 0};
    if (word < nwords) {
      const int j = (int)(word % Wd);
      const long long r = word / Wd;
      const int y = (int)(r % H);
      const long long b = r / H;
      const int x0 = j * 32;
      w.npx = min(32, W - x0);
      w.xoff = b * x_sb + y * x_sy + x0 * x_sx;
      w.soff = b * s_sb + y * s_sy + (long long)x0 * sp;
      w.pix = (b * H + y) * W + x0;
    }
    wi[threadIdx.x] = w;
    s_word[threadIdx.x] = 0u;
  }
  __syncthreads();
}

// phase 1: x -> xs[(w*32 + px) * cps + cc]; chunks beyond C are zero-filled
template <typename T, int VEC>
__device__ __forceinline__ void planar_load_tile(uint4* xs, const PlanarWord* wi, const T* __restrict__ x,
                                                 long long x_sc, long long x_sx, int nw, int cpv, int cps,
                                                 int C, unsigned cpv_magic) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll 2
  for (int p = wid; p < nw * cpv; p += 8) {
    const int w = cpv == 1 ? p : (int)__umulhi((unsigned)p, cpv_magic), cc = p - w * cpv;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    T* ve = reinterpret_cast<T*>(&v);
    if (lane < wi[w].npx) {
      const T* xp = x + wi[w].xoff + lane * x_sx;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int c = cc * VEC + e;
        if (c < C) ve[e] = __ldg(xp + c * x_sc);
      }
    }
    xs[(w * 32 + lane) * cps + cc] = v;
  }
}

// OR the pixel flag of this thread into its word's accumulator (warp-reduced when the warp's 32
// chunks belong to one word, which is the common case)
__device__ __forceinline__ void planar_flag(unsigned* s_word, int w, unsigned m) {
  const int w0 = __shfl_sync(0xffffffffu, w, 0), w1 = __shfl_sync(0xffffffffu, w, 31);
  if (w0 == w1) {
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_word[w0], m);
  } else if (m) {
    atomicOr(&s_word[w], m);
  }
}

template <typename T, int VEC, int UPDATE>
__global__ void __launch_bounds__(256)
detect_planar_kernel(const T* __restrict__ x, long long x_sb, long long x_sc, long long x_sy,
                     T* __restrict__ st, long long s_sb, long long s_sy, int sp, AuxPlanes aux,
                     uint32_t* __restrict__ bits, int B, int H, int W, int C, int Wd, T thr,
                     unsigned cpv_magic, int cps, int nw) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char dp_smem[];
  uint4* xs = reinterpret_cast<uint4*>(dp_smem);                 // [nw*32][cps]
  __shared__ unsigned s_word[kPlanarMaxWords];
  __shared__ PlanarWord wi[kPlanarMaxWords];
  const long long nwords = (long long)B * H * Wd;
  const int word0 = blockIdx.x * nw;
  planar_words_setup(wi, s_word, word0, nw, nwords, H, W, Wd, x_sb, x_sy, 1, s_sb, s_sy, sp);
  const int cpv = (C + VEC - 1) / VEC, tail = C % VEC;
  planar_load_tile<T, VEC>(xs, wi, x, x_sc, 1, nw, cpv, cps, C, cpv_magic);
  __syncthreads();
  const int nq = nw * 32 * cpv;
  for (int q0 = 0; q0 < nq; q0 += 256) {
    const int q = min(q0 + (int)threadIdx.x, nq - 1);            // (clamped lanes take no part below)
    const bool in = q0 + (int)threadIdx.x < nq;
    const int t = cpv == 1 ? q : (int)__umulhi((unsigned)q, cpv_magic), cc = q - t * cpv;
    const int w = t >> 5, px = t & 31;
    unsigned m = 0u;
    if (in && px < wi[w].npx) {
      uint4 xv = xs[t * cps + cc];
      T* sptr = st + wi[w].soff + (long long)px * sp + cc * VEC;
      const uint4 sv = ld16(sptr);
      if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, sv, tail);
      if (Chunk<T>::changed(sv, xv, thr)) m = 1u << px;
      if (UPDATE == CB_UPDATE_ALL) store_state<T>(sptr, xv, aux, wi[w].pix + px, cc * VEC);
    }
    planar_flag(s_word, w, m);
  }
  __syncthreads();
  if (threadIdx.x < (unsigned)nw && wi[threadIdx.x].npx > 0) bits[word0 + threadIdx.x] = s_word[threadIdx.x];
  if (UPDATE == CB_UPDATE_CHANGED) {              // feedback: accept the new value at changed pixels
    for (int q = threadIdx.x; q < nq; q += 256) {
      const int t = cpv == 1 ? q : (int)__umulhi((unsigned)q, cpv_magic), cc = q - t * cpv;
      const int w = t >> 5, px = t & 31;
      if ((s_word[w] >> px) & 1u) {
        uint4 xv = xs[t * cps + cc];
        T* sptr = st + wi[w].soff + (long long)px * sp + cc * VEC;
        if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, ld16(sptr), tail);
        store_state<T>(sptr, xv, aux, wi[w].pix + px, cc * VEC);
      }
    }
  }
}

// words per block of the planar kernels: every warp gets >= 2 (word, chunk) pairs, within 64 KB
inline int planar_words_per_block(int cpv, int cps) {
  static const int forced = [] {
    const char* e = getenv("CBINFER_PLANAR_NW");               // tuning knob: 1, 2, 4 or 8
    return e ? atoi(e) : 0;
  }();
  if (forced > 0 && forced <= kPlanarMaxWords && (forced & (forced - 1)) == 0 &&
      (size_t)forced * 32 * cps * 16 <= 96 * 1024)
    return forced;
  // measured (tools/planar_bench.py, L2 flushed): 4 (word, chunk) pairs per warp, at most 4 words --
  // C=64 fp32 153 -> 135 us at nw=2, bf16 100 -> 94 us at nw=4; 8 words lose again (fewer, fatter blocks)
  int nw = 1;
  while (nw < 4 && nw * cpv < 32 && (size_t)(2 * nw) * 32 * cps * 16 <= 64 * 1024) nw *= 2;
  return nw;
}

template <typename T, int UPDATE>
__global__ void __launch_bounds__(256)
detect_generic_kernel(const T* __restrict__ x, long long x_sb, long long x_sc, long long x_sy,
                      long long x_sx, T* __restrict__ st, long long s_sb, long long s_sc,
                      long long s_sy, long long s_sx, AuxPlanes aux,
                      uint32_t* __restrict__ bits, int B, int H, int W, int C, int Wd, T thr) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  // 32-bit index math (the host checks words < 2^31): two 64-bit divisions per warp were most of this
  // kernel's instructions (ncu: 173 warp instructions per 32 pixels, SM 54 % busy on an HBM-bound scan)
  const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= (unsigned)(B * H * Wd)) return;
  const unsigned r = warp / (unsigned)Wd;
  const int j = (int)(warp - r * (unsigned)Wd);
  const int b = (int)(r / (unsigned)H);
  const int y = (int)(r - (unsigned)b * (unsigned)H);
  const int xx = j * 32 + lane;
  bool f = false;
  const T* xp = x + b * x_sb + y * x_sy + xx * x_sx;
  T* sp = st + b * s_sb + y * s_sy + xx * s_sx;
  const long long gpix = ((long long)b * H + y) * W + xx;
  if (xx < W) {
    for (int c = 0; c < C; ++c) {
      const T xv = xp[c * x_sc];
      const T sv = sp[c * s_sc];
      f |= value_changed(sv, xv, thr);
      if (UPDATE == CB_UPDATE_ALL) store_state_scalar(sp + c * s_sc, xv, aux, gpix, c);
    }
    if (UPDATE == CB_UPDATE_CHANGED && f)
      for (int c = 0; c < C; ++c) store_state_scalar(sp + c * s_sc, xp[c * x_sc], aux, gpix, c);
  }
  const unsigned word = __ballot_sync(0xffffffffu, f);
  if (lane == 0) bits[warp] = word;
}

// uint8 frame ingest (decoder / camera output, any strides: HWC interleaved or planar) against a
// pixel-major fp32 state whose pixel is one 16-byte chunk (C <= 4).  The normalisation of the
// reference's readers -- frame/255 (sceneLabeling/videoSequenceReader.py:65), frame/256 - 0.5
// (openPose/PoseDetector.py:72) -- is applied on the fly with IEEE division and addition, so the
// result is bit-identical to detecting on the host-normalised fp32 frame, at a quarter of the
// host->device bytes.
template <int UPDATE>
__global__ void __launch_bounds__(256)
detect_u8_kernel(const uint8_t* __restrict__ x, long long x_sb, long long x_sc, long long x_sy,
                 long long x_sx, float* __restrict__ st, long long s_sb, long long s_sy,
                 AuxPlanes aux, uint32_t* __restrict__ bits, int B, int H, int W, int C, int Wd,
                 float divisor, float bias, float thr) {
  pdl_prologue();
  // the 256 possible normalised values, computed once per block with the IEEE division and addition the
  // host would use (bit-identical by construction): a table lookup per channel instead of a division
  __shared__ float lut[256];
  lut[threadIdx.x] = __fadd_rn(__fdiv_rn((float)threadIdx.x, divisor), bias);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // 32-bit index math (the host checks words < 2^31): two 64-bit divisions per warp were most of this
  // kernel's instructions (ncu: 173 warp instructions per 32 pixels, SM 54 % busy on an HBM-bound scan)
  const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= (unsigned)(B * H * Wd)) return;
  const unsigned r = warp / (unsigned)Wd;
  const int j = (int)(warp - r * (unsigned)Wd);
  const int b = (int)(r / (unsigned)H);
  const int y = (int)(r - (unsigned)b * (unsigned)H);
  const int xx = j * 32 + lane;
  bool f = false;
  if (xx < W) {
    const uint8_t* xp = x + b * x_sb + y * x_sy + xx * x_sx;
    float* sp = st + b * s_sb + y * s_sy + (long long)xx * 4;
    const uint4 sv = ld16_stream(sp);
    uint4 nv = sv;                             // pad lanes keep the state's (zero) value
    float* ne = reinterpret_cast<float*>(&nv);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) ne[c] = lut[__ldcs(xp + c * x_sc)];
    f = Chunk<float>::changed(sv, nv, thr);
    if (UPDATE == CB_UPDATE_ALL || (UPDATE == CB_UPDATE_CHANGED && f))
      store_state<float>(sp, nv, aux, ((long long)b * H + y) * W + xx, 0);
  }
  const unsigned word = __ballot_sync(0xffffffffu, f);
  if (lane == 0) bits[warp] = word;
}

template <typename T> __host__ __device__ inline T thr_cast(float t);
template <> __host__ __device__ inline float thr_cast<float>(float t) { return t; }
template <> __host__ __device__ inline __half thr_cast<__half>(float t) { return __float2half_rn(t); }
template <> __host__ __device__ inline __nv_bfloat16 thr_cast<__nv_bfloat16>(float t) {
  return __float2bfloat16_rn(t);
}


template <typename T>
inline int make_aux(AuxPlanes& aux, int aux_mode, void* aux_hi, void* aux_lo, const void* state,
                    int C) {
  aux = AuxPlanes{0, 0, 0, nullptr, nullptr};
  if (aux_mode == 0) return 0;
  CB_CHECK_ARG(sizeof(T) == 4, "change_detect: auxiliary operand planes exist for fp32 data only");
  if (aux_mode == 1) {
    CB_CHECK_ARG(aux_lo && ((uintptr_t)aux_lo % 16) == 0, "change_detect: bad tf32 remainder plane");
    aux.mode = 1;
    aux.lo_off = (long long)((const T*)aux_lo - (const T*)state);
    return 0;
  }
  CB_CHECK_ARG(aux_mode == 2, "change_detect: bad aux_mode %d", aux_mode);
  CB_CHECK_ARG(aux_hi && aux_lo && ((uintptr_t)aux_hi % 16) == 0 && ((uintptr_t)aux_lo % 16) == 0,
               "change_detect: bad bf16 operand planes");
  aux.mode = 2;
  aux.pitch16 = pitch16_of(C);
  aux.hi16 = (__nv_bfloat16*)aux_hi;
  aux.lo16 = (__nv_bfloat16*)aux_lo;
  return 0;
}

template <typename T, int VEC>
int launch_detect(cudaStream_t stream, const void* x, long long x_sb, long long x_sc,
                  long long x_sy, long long x_sx, void* state, long long s_sb, long long s_sc,
                  long long s_sy, long long s_sx, int aux_mode, void* aux_hi, void* aux_lo,
                  uint32_t* bits, int B, int C, int H, int W, float threshold, int update) {
  const int Wd = (W + 31) / 32;
  const long long words = (long long)B * H * Wd;
  if (words == 0) return 0;
  const T thr = thr_cast<T>(threshold);
  const size_t es = sizeof(T);
  AuxPlanes aux;
  if (int rc = make_aux<T>(aux, aux_mode, aux_hi, aux_lo, state, C)) return rc;
  const bool vec_ok = x_sc == 1 && s_sc == 1 && (x_sx % VEC) == 0 && (s_sx % VEC) == 0 &&
                      x_sx >= C && s_sx >= C && ((x_sy * es) % 16) == 0 && ((s_sy * es) % 16) == 0 &&
                      ((x_sb * es) % 16) == 0 && ((s_sb * es) % 16) == 0 &&
                      ((uintptr_t)x % 16) == 0 && ((uintptr_t)state % 16) == 0 &&
                      x_sx < (1ll << 30) && s_sx < (1ll << 30);
  const bool narrow_ok = !vec_ok && s_sc == 1 && s_sx == VEC && C <= VEC &&
                         ((s_sy * es) % 16) == 0 && ((s_sb * es) % 16) == 0 &&
                         ((uintptr_t)state % 16) == 0;
  const int cpv = (C + VEC - 1) / VEC;
  const unsigned magic = cpv > 1 ? (unsigned)((0x100000000ull + cpv - 1) / cpv) : 0u;
  // planar x, pixel-major multi-chunk state: transposed through shared memory (detect_planar_kernel)
  const int cps = cpv | 1;
  const int planar_nw = planar_words_per_block(cpv, cps);
  const size_t planar_smem = (size_t)planar_nw * 32 * cps * 16;
  const bool planar_ok = !vec_ok && !narrow_ok && x_sx == 1 && s_sc == 1 && (s_sx % VEC) == 0 && s_sx >= C &&
                         ((s_sy * es) % 16) == 0 && ((s_sb * es) % 16) == 0 && ((uintptr_t)state % 16) == 0 &&
                         s_sx < (1ll << 30) && planar_smem <= 200 * 1024;
  // warps per word: keep >= 4 load batches (of U*32 chunks) per warp, at most one block per word
  int wlog = 0;
  while (wlog < 3 && (32 * cpv) / (1 << (wlog + 1)) >= 4 * 4 * 32) ++wlog;
  const long long vec_blocks = (words + (kDetWarps >> wlog) - 1) / (kDetWarps >> wlog);
  const long long blocks = vec_ok ? vec_blocks : (words + 7) / 8;
  CB_CHECK_ARG(blocks < (1ll << 31) && words < (1ll << 31), "change_detect: image too large");
  dim3 grid((unsigned)blocks), block(256);
#define CB_DET(U_)                                                                             \
  if (vec_ok) {                                                                                \
    if (cpv >= 4)                                                                              \
      cb::launch_pdl(detect_vec_kernel<T, VEC, U_, 4>, grid, block, 0, stream,                             \
          (const T*)x, x_sb, x_sy, (int)x_sx, (T*)state, s_sb, s_sy, (int)s_sx, aux, bits,  \
          B, H, W, C, Wd, thr, magic, wlog);                                                   \
    else                                                                                       \
      cb::launch_pdl(detect_vec_kernel<T, VEC, U_, 2>, grid, block, 0, stream,                             \
          (const T*)x, x_sb, x_sy, (int)x_sx, (T*)state, s_sb, s_sy, (int)s_sx, aux, bits,  \
          B, H, W, C, Wd, thr, magic, wlog);                                                   \
  } else if (narrow_ok) {                                                                      \
    cb::launch_pdl(detect_narrow_kernel<T, VEC, U_>, grid, block, 0, stream,                               \
        (const T*)x, x_sb, x_sc, x_sy, x_sx, (T*)state, s_sb, s_sy, aux, bits, B, H, W, C,  \
        Wd, thr);                                                                              \
  } else if (planar_ok) {                                                                      \
    if (planar_smem > 48 * 1024)                                                               \
      cudaFuncSetAttribute(detect_planar_kernel<T, VEC, U_>,                                   \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)planar_smem);     \
    cb::launch_pdl(detect_planar_kernel<T, VEC, U_>, dim3((unsigned)((words + planar_nw - 1) / planar_nw)), \
        block, planar_smem, stream,                                                            \
        (const T*)x, x_sb, x_sc, x_sy, (T*)state, s_sb, s_sy, (int)s_sx, aux, bits, B, H, W, C, \
        Wd, thr, magic, cps, planar_nw);                                                       \
  } else {                                                                                     \
    cb::launch_pdl(detect_generic_kernel<T, U_>, grid, block, 0, stream,                                   \
        (const T*)x, x_sb, x_sc, x_sy, x_sx, (T*)state, s_sb, s_sc, s_sy, s_sx, aux, bits,  \
        B, H, W, C, Wd, thr);                                                                  \
  }
  switch (update) {
    case CB_UPDATE_NONE: CB_DET(CB_UPDATE_NONE) break;
    case CB_UPDATE_CHANGED: CB_DET(CB_UPDATE_CHANGED) break;
    case CB_UPDATE_ALL: CB_DET(CB_UPDATE_ALL) break;
    default: return fail(2, "change_detect: bad update_mode %d", update);
  }
#undef CB_DET
  CB_CHECK_LAUNCH("change_detect");
  return 0;
}


inline int launch_detect_u8(cudaStream_t stream, const void* x, long long x_sb, long long x_sc,
                            long long x_sy, long long x_sx, void* state, long long s_sb,
                            long long s_sc, long long s_sy, long long s_sx, int aux_mode,
                            void* aux_hi, void* aux_lo, uint32_t* bits, int B, int C, int H, int W,
                            float divisor, float bias, float threshold, int update) {
  const int Wd = (W + 31) / 32;
  const long long words = (long long)B * H * Wd;
  if (words == 0) return 0;
  CB_CHECK_ARG(C <= 4 && s_sc == 1 && s_sx == 4 && (s_sy % 4) == 0 && (s_sb % 4) == 0 &&
                   ((uintptr_t)state % 16) == 0,
               "change_detect_u8: needs <= 4 channels and a pixel-major fp32 state of pitch 4");
  CB_CHECK_ARG(divisor != 0.f, "change_detect_u8: divisor must be non-zero");
  AuxPlanes aux;
  if (int rc = make_aux<float>(aux, aux_mode, aux_hi, aux_lo, state, C)) return rc;
  const long long blocks = (words + 7) / 8;
  CB_CHECK_ARG(blocks < (1ll << 31), "change_detect_u8: image too large");
#define CB_DET8(U_)                                                                              \
  cb::launch_pdl(detect_u8_kernel<U_>, dim3((unsigned)blocks), dim3(256), 0, stream,            \
                 (const uint8_t*)x, x_sb, x_sc, x_sy, x_sx, (float*)state, s_sb, s_sy, aux, bits, \
                 B, H, W, C, Wd, divisor, bias, threshold);
  switch (update) {
    case CB_UPDATE_NONE: CB_DET8(CB_UPDATE_NONE) break;
    case CB_UPDATE_CHANGED: CB_DET8(CB_UPDATE_CHANGED) break;
    case CB_UPDATE_ALL: CB_DET8(CB_UPDATE_ALL) break;
    default: return fail(2, "change_detect_u8: bad update_mode %d", update);
  }
#undef CB_DET8
  CB_CHECK_LAUNCH("change_detect_u8");
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Candidate ("sparse") detection: the same per-pixel test, evaluated only at the pixels an
// upstream change-based layer reports as rewritten.  Exact w.r.t. the dense scan as long as x is
// unchanged everywhere else, the threshold was not lowered and the state is not fresh (the module
// checks this): an untouched pixel was either accepted last frame (state == x, difference 0) or
// is as far below the threshold as it was.  Replaces the O(C*P) scan with O(C*n_candidates).
// A group of 1 << glog lanes owns one candidate (16-byte chunks, pixel-major); set bits are
// OR-ed into the pre-zeroed bitmap.
// ------------------------------------------------------------------------------------------------
// FUSE: the last block to finish also dilates and compacts the (small) bitmap, see block_dilate_compact
struct FusedCompact {
  uint32_t* dil_bits;
  int32_t* idx;
  int32_t* count;
  unsigned* sync;                // one word, zero at rest: blocks done
  int kh, kw, nwords, clear;
};

template <typename T, int VEC, int UPDATE, bool FUSE = false>
__global__ void __launch_bounds__(256)
detect_sparse_vec_kernel(const T* __restrict__ x, long long x_sb, long long x_sy, int xp,
                         T* __restrict__ st, long long s_sb, long long s_sy, int sp,
                         AuxPlanes aux, const int32_t* __restrict__ cand,
                         const int32_t* __restrict__ ncand, uint32_t* __restrict__ bits, int H,
                         int W, int C, int Wd, T thr, int glog, FusedCompact fc) {
  pdl_prologue();
  const int n = *ncand;
  const int lane = threadIdx.x & 31;
  const int G = 1 << glog, ppw = 32 >> glog;
  const int sub = lane >> glog, gl = lane & (G - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (sub * G);
  const int P = H * W;
  const int cpv = (C + VEC - 1) / VEC, tail = C % VEC;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long j0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ppw;
       j0 < n; j0 += nwarps * ppw) {
    const long long j = j0 + sub;
    const bool have = j < n;
    int b = 0, y = 0, xx = 0, gpix = 0;
    if (have) {
      const int pix = __ldg(cand + j);
      gpix = pix;
      b = pix / P;
      const int p = pix - b * P;
      y = p / W;
      xx = p - y * W;
    }
    const T* xb = x + b * x_sb + y * x_sy + (long long)xx * xp;
    T* sb = st + b * s_sb + y * s_sy + (long long)xx * sp;
    bool f = false;
    if (have) {
      for (int cc = gl; cc < cpv; cc += G) {
        uint4 xv = ldg16(xb + cc * VEC);
        const uint4 sv = ld16(sb + cc * VEC);
        if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, sv, tail);
        f |= Chunk<T>::changed(sv, xv, thr);
        if (UPDATE == CB_UPDATE_ALL) store_state<T>(sb + cc * VEC, xv, aux, gpix, cc * VEC);
      }
    }
    const bool chg = (__ballot_sync(0xffffffffu, f) & gmask) != 0u;   // any lane of my group
    if (have && chg) {
      if (gl == 0) atomicOr(bits + ((long long)b * H + y) * Wd + (xx >> 5), 1u << (xx & 31));
      if (UPDATE == CB_UPDATE_CHANGED) {
        for (int cc = gl; cc < cpv; cc += G) {
          uint4 xv = ldg16(xb + cc * VEC);
          if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, ld16(sb + cc * VEC), tail);
          store_state<T>(sb + cc * VEC, xv, aux, gpix, cc * VEC);
        }
      }
    }
  }
  if (FUSE) {
    __shared__ uint32_t s_win[kCompactWin];
    __shared__ int s_warp[8];
    __shared__ unsigned s_last;
    __threadfence();                                          // my bitmap bits precede my done-increment
    __syncthreads();
    if (threadIdx.x == 0) {
      s_last = atomicAdd(fc.sync, 1u) == gridDim.x - 1u;
      if (s_last) *fc.sync = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    block_dilate_compact(bits, fc.dil_bits, fc.idx, fc.count, s_win, s_warp, H, W, Wd, fc.kh, fc.kw, fc.nwords,
                         fc.clear != 0);
  }
}

template <typename T, int UPDATE>
__global__ void __launch_bounds__(256)
detect_sparse_generic_kernel(const T* __restrict__ x, long long x_sb, long long x_sc,
                             long long x_sy, long long x_sx, T* __restrict__ st, long long s_sb,
                             long long s_sc, long long s_sy, long long s_sx, AuxPlanes aux,
                             const int32_t* __restrict__ cand, const int32_t* __restrict__ ncand,
                             uint32_t* __restrict__ bits, int H, int W, int C, int Wd, T thr) {
  pdl_prologue();
  const int n = *ncand;
  const int P = H * W;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const int pix = __ldg(cand + j);
    const int b = pix / P, p = pix - b * P;
    const int y = p / W, xx = p - y * W;
    const T* xp = x + b * x_sb + y * x_sy + xx * x_sx;
    T* sp = st + b * s_sb + y * s_sy + xx * s_sx;
    bool f = false;
    for (int c = 0; c < C; ++c) {
      const T xv = xp[c * x_sc];
      f |= value_changed(sp[c * s_sc], xv, thr);
      if (UPDATE == CB_UPDATE_ALL) store_state_scalar(sp + c * s_sc, xv, aux, (long long)pix, c);
    }
    if (f) {
      atomicOr(bits + ((long long)b * H + y) * Wd + (xx >> 5), 1u << (xx & 31));
      if (UPDATE == CB_UPDATE_CHANGED)
        for (int c = 0; c < C; ++c) store_state_scalar(sp + c * s_sc, xp[c * x_sc], aux, (long long)pix, c);
    }
  }
}

template <typename T, int VEC>
int launch_detect_sparse(cudaStream_t stream, const void* x, long long x_sb, long long x_sc,
                         long long x_sy, long long x_sx, void* state, long long s_sb,
                         long long s_sc, long long s_sy, long long s_sx, int aux_mode,
                         void* aux_hi, void* aux_lo, const int32_t* cand, const int32_t* ncand, uint32_t* bits, int B, int C,
                         int H, int W, float threshold, int update, int bits_are_clear,
                         const FusedCompact* fuse = nullptr) {
  const int Wd = (W + 31) / 32;
  const long long words = (long long)B * H * Wd;
  if (words == 0) return 0;
  if (!bits_are_clear && cudaMemsetAsync(bits, 0, (size_t)words * 4, stream) != cudaSuccess)
    return fail(3, "change_detect_sparse: memset failed");
  const T thr = thr_cast<T>(threshold);
  const size_t es = sizeof(T);
  AuxPlanes aux;
  if (int rc = make_aux<T>(aux, aux_mode, aux_hi, aux_lo, state, C)) return rc;
  const bool vec_ok = x_sc == 1 && s_sc == 1 && (x_sx % VEC) == 0 && (s_sx % VEC) == 0 &&
                      x_sx >= C && s_sx >= C && ((x_sy * es) % 16) == 0 && ((s_sy * es) % 16) == 0 &&
                      ((x_sb * es) % 16) == 0 && ((s_sb * es) % 16) == 0 &&
                      ((uintptr_t)x % 16) == 0 && ((uintptr_t)state % 16) == 0 &&
                      x_sx < (1ll << 30) && s_sx < (1ll << 30);
  unsigned grid = (unsigned)(sm_count() * 16);
  const int cpv = (C + VEC - 1) / VEC;
  int glog = 0;
  while ((1 << glog) < cpv && glog < 5) ++glog;
  glog = cb::glog_tuned(glog);
  {
    // small maps: no more blocks than there are pixels to look at (every block of the fused variant
    // passes through one atomic)
    const long long ppb = (long long)(32 >> glog) * 8;         // candidates per block and round
    const long long need = ((long long)B * H * W + ppb - 1) / ppb;
    if (need < (long long)grid) grid = (unsigned)(need < 1 ? 1 : need);
  }
  CB_CHECK_ARG(!fuse || (vec_ok && words <= kCompactWin), "change_detect_sparse_compact: needs pixel-major "
               "tensors and a bitmap of at most %d words", kCompactWin);
  FusedCompact fc = fuse ? *fuse : FusedCompact{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
#define CB_DETS(U_)                                                                              \
  if (vec_ok && fuse)                                                                            \
    cb::launch_pdl(detect_sparse_vec_kernel<T, VEC, U_, true>, grid, 256, 0, stream,                         \
        (const T*)x, x_sb, x_sy, (int)x_sx, (T*)state, s_sb, s_sy, (int)s_sx, aux, cand,      \
        ncand, bits, H, W, C, Wd, thr, glog, fc);                                                \
  else if (vec_ok)                                                                               \
    cb::launch_pdl(detect_sparse_vec_kernel<T, VEC, U_, false>, grid, 256, 0, stream,                        \
        (const T*)x, x_sb, x_sy, (int)x_sx, (T*)state, s_sb, s_sy, (int)s_sx, aux, cand,      \
        ncand, bits, H, W, C, Wd, thr, glog, fc);                                                \
  else                                                                                           \
    cb::launch_pdl(detect_sparse_generic_kernel<T, U_>, grid, 256, 0, stream,                                \
        (const T*)x, x_sb, x_sc, x_sy, x_sx, (T*)state, s_sb, s_sc, s_sy, s_sx, aux, cand,    \
        ncand, bits, H, W, C, Wd, thr);
  switch (update) {
    case CB_UPDATE_NONE: CB_DETS(CB_UPDATE_NONE) break;
    case CB_UPDATE_CHANGED: CB_DETS(CB_UPDATE_CHANGED) break;
    case CB_UPDATE_ALL: CB_DETS(CB_UPDATE_ALL) break;
    default: return fail(2, "change_detect_sparse: bad update_mode %d", update);
  }
#undef CB_DETS
  CB_CHECK_LAUNCH("change_detect_sparse");
  return 0;
}

}  // namespace cb

namespace cb {

// ------------------------------------------------------------------------------------------------
// Candidate detection fused with ordered compaction, for layers whose change set needs no dilation
// (1x1 kernels): one launch instead of memset + sparse detection + dilate/compact.
//   phase 1 (all blocks): per candidate, threshold + state maintenance, flag byte per candidate;
//   phase 2 (the last block to finish): ordered compaction of the flagged candidates -> idx, count
//           (and, optionally, their bits in a pre-cleared bitmap).
// ws: [0] = blocks-done counter (left at 0), then one flag byte per candidate.
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int UPDATE>
__global__ void __launch_bounds__(256)
detect_compact_sparse_kernel(const T* __restrict__ x, long long x_sb, long long x_sy, int xp,
                             T* __restrict__ st, long long s_sb, long long s_sy, int sp,
                             AuxPlanes aux, const int32_t* __restrict__ cand,
                             const int32_t* __restrict__ ncand, int32_t* __restrict__ idx,
                             int32_t* __restrict__ count, uint32_t* __restrict__ bits,
                             unsigned* __restrict__ ws, int H, int W, int C, int Wd, T thr,
                             int glog) {
  pdl_prologue();
  const int n = *ncand;
  uint8_t* flags = reinterpret_cast<uint8_t*>(ws + 4);
  const int lane = threadIdx.x & 31;
  const int G = 1 << glog, ppw = 32 >> glog;
  const int sub = lane >> glog, gl = lane & (G - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (sub * G);
  const int P = H * W;
  const int cpv = (C + VEC - 1) / VEC, tail = C % VEC;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long j0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ppw;
       j0 < n; j0 += nwarps * ppw) {
    const long long j = j0 + sub;
    const bool have = j < n;
    int b = 0, y = 0, xx = 0, gpix = 0;
    if (have) {
      gpix = cand[j];
      b = gpix / P;
      const int p = gpix - b * P;
      y = p / W;
      xx = p - y * W;
    }
    const T* xb = x + b * x_sb + y * x_sy + (long long)xx * xp;
    T* sb = st + b * s_sb + y * s_sy + (long long)xx * sp;
    bool f = false;
    if (have) {
      for (int cc = gl; cc < cpv; cc += G) {
        uint4 xv = ld16(xb + cc * VEC);
        const uint4 sv = ld16(sb + cc * VEC);
        if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, sv, tail);
        f |= Chunk<T>::changed(sv, xv, thr);
        if (UPDATE == CB_UPDATE_ALL) store_state<T>(sb + cc * VEC, xv, aux, gpix, cc * VEC);
      }
    }
    const bool chg = (__ballot_sync(0xffffffffu, f) & gmask) != 0u;
    if (have) {
      if (gl == 0) flags[j] = chg ? 1 : 0;
      if (chg && UPDATE == CB_UPDATE_CHANGED) {
        for (int cc = gl; cc < cpv; cc += G) {
          uint4 xv = ld16(xb + cc * VEC);
          if (tail && cc == cpv - 1) xv = merge_tail<T, VEC>(xv, ld16(sb + cc * VEC), tail);
          store_state<T>(sb + cc * VEC, xv, aux, gpix, cc * VEC);
        }
      }
    }
  }
  // ---- the last block compacts ---------------------------------------------------------------
  __shared__ unsigned s_last;
  __shared__ int s_wsum[8];
  __shared__ int s_run;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(ws, 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) ws[0] = 0;
    s_run = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int wid = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 256 * 8) {          // 8 candidates per thread per round
    const int j0 = base + threadIdx.x * 8;
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (j0 + e < n && reinterpret_cast<volatile uint8_t*>(flags)[j0 + e]) m |= 1u << e;
    const int cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int v = s_wsum[q];
      if (q < wid) woff += v;
      total += v;
    }
    int o = s_run + woff + incl - cnt;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (m & (1u << e)) {
        const int pix = cand[j0 + e];
        idx[o++] = pix;
        if (bits) {
          const int b = pix / P, p = pix - b * P;
          const int y = p / W, xx = p - y * W;
          atomicOr(bits + ((long long)b * H + y) * Wd + (xx >> 5), 1u << (xx & 31));
        }
      }
    __syncthreads();
    if (threadIdx.x == 0) s_run += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_run;
}

template <typename T, int VEC>
int launch_detect_compact_sparse(cudaStream_t stream, const void* x, long long x_sb, long long x_sc,
                                 long long x_sy, long long x_sx, void* state, long long s_sb,
                                 long long s_sc, long long s_sy, long long s_sx, int aux_mode,
                                 void* aux_hi, void* aux_lo, const int32_t* cand,
                                 const int32_t* ncand, int32_t* idx, int32_t* count, uint32_t* bits,
                                 void* ws, int B, int C, int H, int W, float threshold, int update) {
  const size_t es = sizeof(T);
  AuxPlanes aux;
  if (int rc = make_aux<T>(aux, aux_mode, aux_hi, aux_lo, state, C)) return rc;
  const bool vec_ok = x_sc == 1 && s_sc == 1 && (x_sx % VEC) == 0 && (s_sx % VEC) == 0 &&
                      x_sx >= C && s_sx >= C && ((x_sy * es) % 16) == 0 && ((s_sy * es) % 16) == 0 &&
                      ((x_sb * es) % 16) == 0 && ((s_sb * es) % 16) == 0 &&
                      ((uintptr_t)x % 16) == 0 && ((uintptr_t)state % 16) == 0 &&
                      x_sx < (1ll << 30) && s_sx < (1ll << 30);
  CB_CHECK_ARG(vec_ok, "detect_compact_sparse: needs pixel-major, 16-byte aligned x and state");
  if ((long long)B * H * W == 0) {
    cudaMemsetAsync(count, 0, sizeof(int32_t), stream);
    return 0;
  }
  const int cpv = (C + VEC - 1) / VEC;
  int glog = 0;
  while ((1 << glog) < cpv && glog < 5) ++glog;
  glog = cb::glog_tuned(glog);
  const unsigned grid = (unsigned)(sm_count() * 8);
  const int Wd = (W + 31) / 32;
  const T thr = thr_cast<T>(threshold);
#define CB_DCS(U_)                                                                                 \
  cb::launch_pdl(detect_compact_sparse_kernel<T, VEC, U_>, grid, 256, 0, stream, (const T*)x,     \
                 x_sb, x_sy, (int)x_sx, (T*)state, s_sb, s_sy, (int)s_sx, aux, cand, ncand, idx,   \
                 count, bits, (unsigned*)ws, H, W, C, Wd, thr, glog);
  switch (update) {
    case CB_UPDATE_NONE: CB_DCS(CB_UPDATE_NONE) break;
    case CB_UPDATE_CHANGED: CB_DCS(CB_UPDATE_CHANGED) break;
    case CB_UPDATE_ALL: CB_DCS(CB_UPDATE_ALL) break;
    default: return fail(2, "detect_compact_sparse: bad update_mode %d", update);
  }
#undef CB_DCS
  CB_CHECK_LAUNCH("detect_compact_sparse");
  return 0;
}

}  // namespace cb
