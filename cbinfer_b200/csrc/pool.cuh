// pool.cuh -- change-based 2x2 / stride-2 max pooling.
//
// Replaces maxPool2d_kernel (reference cbconv2d_cg_backend.cu:199-227, half :207-237): for each
// changed *input* pixel recompute its window maximum over all channels (init -inf, window clipped
// to the input) into the persistent pooled map.  One warp per changed pixel, lanes over channels
// (contiguous for the pixel-major layout).  When the change bitmap is supplied each window is
// recomputed once (by its first changed pixel) instead of up to four times; windows whose output
// coordinate is outside [oH,oW) are skipped (the reference writes out of bounds there).
// Traffic per distinct window: 4*C*s read + C*s written.
#pragma once
#include "cb_common.cuh"

namespace cb {

__device__ __forceinline__ bool bit_at(const uint32_t* bits, long long row, int Wd, int x) {
  return (__ldg(bits + row * Wd + (x >> 5)) >> (x & 31)) & 1u;
}

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ __half neg_inf<__half>() { return __ushort_as_half(0xFC00); }
template <> __device__ __forceinline__ __nv_bfloat16 neg_inf<__nv_bfloat16>() {
  return __ushort_as_bfloat16(0xFF80);
}
__device__ __forceinline__ float max_keep(float v, float u) { return fmaxf(v, u); }   // cg.cu:220
__device__ __forceinline__ __half max_keep(__half v, __half u) { return __hgt(u, v) ? u : v; }  // half.cu:229
__device__ __forceinline__ __nv_bfloat16 max_keep(__nv_bfloat16 v, __nv_bfloat16 u) {
  return __hgt(u, v) ? u : v;
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const T* __restrict__ x, long long x_sb, long long x_sc, long long x_sy,
                  long long x_sx, const int32_t* __restrict__ idx,
                  const int32_t* __restrict__ count, const uint32_t* __restrict__ bits,
                  T* __restrict__ out, long long o_sb, long long o_sc, long long o_sy,
                  long long o_sx, int C, int H, int W, int oH, int oW) {
  const int n = *count;
  const int lane = threadIdx.x & 31;
  const int P = H * W, Wd = (W + 31) >> 5;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += wstride) {
    const int pix = idx[j];
    const int b = pix / P, p = pix - b * P;
    const int y = p / W, xx = p - y * W;
    const int yo = y >> 1, xo = xx >> 1;
    if (yo >= oH || xo >= oW) continue;
    if (bits) {                                // first changed pixel of the window owns it
      const long long r = (long long)b * H + y;
      bool owner = true;
      if (xx & 1) owner = !bit_at(bits, r, Wd, xx - 1);
      if (owner && (y & 1)) {
        const int xe = xx & ~1;
        owner = !bit_at(bits, r - 1, Wd, xe) && !(xe + 1 < W && bit_at(bits, r - 1, Wd, xe + 1));
      }
      if (!owner) continue;
    }
    const int y0 = yo * 2, x0 = xo * 2;
    const bool hy = y0 + 1 < H, hx = x0 + 1 < W;
    const T* base = x + b * x_sb + y0 * x_sy + x0 * x_sx;
    T* o = out + b * o_sb + yo * o_sy + xo * o_sx;
    for (int c = lane; c < C; c += 32) {
      const T* q = base + c * x_sc;
      T v = neg_inf<T>();
      v = max_keep(v, q[0]);
      if (hx) v = max_keep(v, q[x_sx]);
      if (hy) {
        v = max_keep(v, q[x_sy]);
        if (hx) v = max_keep(v, q[x_sy + x_sx]);
      }
      o[c * o_sc] = v;
    }
  }
}

}  // namespace cb
