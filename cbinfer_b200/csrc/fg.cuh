// fg.cuh -- fine-grained change-based convolution (fp32, planar layout as in the reference).
//
// Replaces changeDetectionFG_kernel (reference cbconv2d_fg_backend.cu:7-23), the per-value
// torch.nonzero (conv2d_fg.py:82) and updateOutputFG_kernel (cbconv2d_fg_backend.cu:37-66),
// fused into one pass: a warp scans 32 input values, ballots the changed ones, then the whole
// warp cooperates on each changed value -- lanes spread over (co, ky, kx) -- and pushes
// W[co,ci,ky,kx]*d into the output with fire-and-forget red.global.add.f32.  The per-value mask
// uses the CUDA reference's strict '>' with flush-to-zero.  prev <- x afterwards (conv2d.py:175).
#pragma once
#include "cb_common.cuh"

namespace cb {

__global__ void __launch_bounds__(256)
fg_update_kernel(const float* __restrict__ x, float* __restrict__ prev,
                 const float* __restrict__ w, float* __restrict__ out, int32_t* __restrict__ count,
                 int B, int Cin, int Cout, int H, int W, int kH, int kW, float thr) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * Cin * H * W;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int taps = kH * kW, work = Cout * taps;
  const long long HW = (long long)H * W;
  for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
       base < total; base += nwarps * 32) {
    const long long i = base + lane;
    float d = 0.f;
    bool m = false;
    if (i < total) {
      const float xv = x[i], pv = prev[i];
      asm("sub.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(xv), "f"(pv));
      m = value_changed(xv, pv, thr);
      prev[i] = xv;
    }
    unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0 && bal) atomicAdd(count, __popc(bal));
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      const float dd = __shfl_sync(0xffffffffu, d, src);
      const long long pos = base + src;
      const int xx = (int)(pos % W);
      const int y = (int)((pos / W) % H);
      const int ci = (int)((pos / HW) % Cin);
      const int b = (int)(pos / (HW * Cin));
      float* ob = out + (long long)b * Cout * HW;
      for (int t = lane; t < work; t += 32) {
        const int co = t / taps, tap = t - co * taps;
        const int ky = tap / kW, kx = tap - ky * kW;
        const int yt = y - ky + kH / 2, xt = xx - kx + kW / 2;
        if (yt >= 0 && yt < H && xt >= 0 && xt < W)
          atomicAdd(ob + ((long long)co * H + yt) * W + xt,
                    __ldg(w + (((long long)co * Cin + ci) * kH + ky) * kW + kx) * dd);
      }
    }
  }
}

}  // namespace cb
