// compact.cuh -- change propagation (dilation) + ordered compaction, bitmap domain.
//
// Replaces the scatter-dilate of changeDetection_kernel (reference cbconv2d_cg_backend.cu:62-72),
// changePropagation_kernel (:101-124) and torch.nonzero(...).int() incl. its host sync
// (conv2d_cg.py:200-213).
//
// One thread owns one bitmap word (32 pixels): vertical OR of the raw rows in the window,
// horizontal dilation with funnel shifts across the neighbouring words, __popc for the count,
// a block scan plus a single-pass chained ("decoupled look-back") scan across tiles for the
// global offsets, then the set bits are expanded to ascending int32 pixel indices.
// Traffic: P/8 bitmap bytes (re-read (2kH+1)*3 times out of L1/L2) + 4n index bytes.
#pragma once
#include "cb_common.cuh"

namespace cb {

#ifndef CB_COMPACT_THREADS
#define CB_COMPACT_THREADS 512
#endif
#ifndef CB_COMPACT_WPT
#define CB_COMPACT_WPT 1
#endif
constexpr int kCompactThreads = CB_COMPACT_THREADS;
constexpr int kCompactWPT = CB_COMPACT_WPT;                  // words per thread
constexpr int kCompactTile = kCompactThreads * kCompactWPT;  // words per tile (16384 pixels; measured best of 256..2048)
constexpr int kCompactWin = 6144;                            // staged raw window (words, 24 KB)

constexpr int kCompactResidentTiles = 2 * 132;  // blocks that are co-resident on any B200 (>= 132 SMs, >= 2 per SM)
struct CompactHeader {                         // first 16 bytes of the workspace
  unsigned reserved, done, epoch, ticket;
};

__host__ __device__ inline size_t compact_ws_bytes(size_t nwords) {
  const size_t tiles = (nwords + kCompactTile - 1) / kCompactTile;
  return sizeof(CompactHeader) + 8 * (tiles + 1);
}

// tile_state word: [63:34] tag (epoch-derived, never 0) | [33:32] status | [31:0] value
__device__ __forceinline__ unsigned long long pack_state(unsigned tag, unsigned status, unsigned v) {
  return ((unsigned long long)tag << 34) | ((unsigned long long)status << 32) | v;
}

// One dilated bitmap word: vertical OR of the raw rows in the window, then horizontal dilation
// with funnel shifts across the neighbouring words.
// `win` holds raw words [win0, win0 + len) of the flat bitmap (staged in shared memory when the
// window fits, else the global bitmap itself with win0 = 0).
__device__ __forceinline__ unsigned dilated_word(const uint32_t* __restrict__ win, int win0, int r,
                                                 int y, int j, int H, int W, int Wd, int kh,
                                                 int kw) {
  unsigned vp = 0, vc = 0, vn = 0;
  const int y0 = max(0, y - kh), y1 = min(H - 1, y + kh);
  const uint32_t* row = win + ((r - y + y0) * Wd + j - win0);
  for (int yy = y0; yy <= y1; ++yy, row += Wd) {
    vc |= row[0];
    if (j > 0) vp |= row[-1];
    if (j + 1 < Wd) vn |= row[1];
  }
  unsigned d = vc;
  for (int dx = 1; dx <= kw; ++dx) {
    d |= (vc << dx) | (vp >> (32 - dx));       // source pixel dx to the left
    d |= (vc >> dx) | (vn << (32 - dx));       // source pixel dx to the right
  }
  if (j == Wd - 1 && (W & 31)) d &= (1u << (W & 31)) - 1u;
  return d;
}

// Word j of row (b, yo) of the 2x2/stride-2 pooled view of a bitmap with Hin x Win pixels:
// bit xo = OR of the four input bits of window (yo, xo).  Used to hand pooled-resolution change
// candidates from a CBPoolMax2d to the next layer.
__device__ __forceinline__ unsigned compress_pairs(unsigned v) {
  unsigned t = (v | (v >> 1)) & 0x55555555u;
  t = (t | (t >> 1)) & 0x33333333u;
  t = (t | (t >> 2)) & 0x0f0f0f0fu;
  t = (t | (t >> 4)) & 0x00ff00ffu;
  t = (t | (t >> 8)) & 0x0000ffffu;
  return t;
}
__device__ __forceinline__ unsigned pooled_word(const uint32_t* __restrict__ raw, int b, int yo,
                                                int j, int Hin, int Wdin, int oW) {
  unsigned v0 = 0, v1 = 0;
  for (int dy = 0; dy < 2; ++dy) {
    const int y = 2 * yo + dy;
    if (y < Hin) {
      const uint32_t* row = raw + ((long long)b * Hin + y) * Wdin;
      if (2 * j < Wdin) v0 |= __ldg(row + 2 * j);
      if (2 * j + 1 < Wdin) v1 |= __ldg(row + 2 * j + 1);
    }
  }
  unsigned d = compress_pairs(v0) | (compress_pairs(v1) << 16);
  const int rem = oW - j * 32;
  if (rem < 32) d &= (1u << rem) - 1u;
  return d;
}

// L2 prefetch hints (extension, no reference counterpart).  The kernels that consume a change set a
// few launches later read rows they have not touched for a frame (the next layer's state at the
// rewritten pixels: cold in L2, a chain of dependent DRAM round trips inside a latency-bound kernel).
// The dilation is the first kernel to know WHICH pixels those are, so it asks L2 for them while the
// contraction in between keeps the tensor cores busy: per dilated word, one cp.async.bulk.prefetch.L2
// per run of set bits and target (pixel-major maps: a run of pixels is one contiguous byte range).
// shift = 1: the target has the 2x2-pooled resolution (issued from even rows).  Hints only -- no
// architectural state changes, results are bit-identical with and without them.
constexpr int kMaxHints = 3;
struct PrefetchHints {
  const char* base[kMaxHints];
  int row_bytes[kMaxHints], shift[kMaxHints], tH[kMaxHints], tW[kMaxHints];
  int n;
};

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void prefetch_hinted_rows(const PrefetchHints& hints, unsigned d, int b, int y, int j) {
#pragma unroll
  for (int h = 0; h < kMaxHints; ++h) {        // (constant indices: the parameter struct stays in constant memory)
    if (h >= hints.n) break;
    const int sh = hints.shift[h];
    if (sh && (y & 1)) continue;
    const int yy = y >> sh, tH = hints.tH[h], tW = hints.tW[h], rb = hints.row_bytes[h];
    if (yy >= tH) continue;
    const char* row = hints.base[h] + (long long)(b * tH + yy) * tW * rb;
    unsigned dd = d;
    while (dd) {
      const int x0 = __ffs(dd) - 1;
      const unsigned inv = ~(dd >> x0);
      const int len = inv ? __ffs(inv) - 1 : 32 - x0;        // run of set bits x0 .. x0 + len - 1
      const int xa = (j * 32 + x0) >> sh;
      int xb = (j * 32 + x0 + len - 1) >> sh;
      if (xb >= tW) xb = tW - 1;
      if (xa <= xb) prefetch_l2_bulk(row + (long long)xa * rb, (unsigned)((xb - xa + 1) * rb));
      dd = x0 + len >= 32 ? 0u : dd & (0xffffffffu << (x0 + len));
    }
  }
}

// List mode: tiles are taken by ticket (atomicAdd on the header's `ticket`), not by blockIdx -- every
// predecessor of a running tile then holds an earlier ticket, i.e. is running or finished, whatever
// order the hardware dispatches blocks in, so the look-back cannot starve (as in CUB's single-pass
// scan).  Tile-only mode has no scan across tiles and keeps blockIdx, and so do grids small enough to
// be resident as a whole.
template <bool HINTS>
__global__ void __launch_bounds__(kCompactThreads)
dilate_compact_kernel(const uint32_t* __restrict__ raw, uint32_t* __restrict__ dil_bits,
                      int8_t* __restrict__ dil_map, int32_t* __restrict__ idx,
                      int32_t* __restrict__ count, void* ws, int B, int H, int W, int Wd, int kh,
                      int kw, int nwords, int ntiles, int pool_hin, int pool_wdin,
                      uint32_t* __restrict__ clear_bits, int32_t* __restrict__ tile_ws, int tile_ty,
                      int tile_xp, int coop, int no_list, const PrefetchHints hints) {
  pdl_prologue();
  CompactHeader* hdr = reinterpret_cast<CompactHeader*>(ws);
  volatile unsigned long long* tstate =
      reinterpret_cast<volatile unsigned long long*>(reinterpret_cast<char*>(ws) + sizeof(CompactHeader));
  constexpr int NW = kCompactThreads / 32;
  __shared__ int s_warp[NW];
  __shared__ int s_base;
  __shared__ uint32_t s_win[kCompactWin];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int tile = blockIdx.x;
  // (up to 2 blocks per SM the whole grid is resident at once and no order can starve anybody: those
  //  launches skip the ticket -- one more dependent global round trip, 2 us inside a step)
  if (!no_list && ntiles > kCompactResidentTiles) {
    if (tid == 0) s_base = (int)atomicAdd(&hdr->ticket, 1u);
    __syncthreads();
    tile = s_base;
    __syncthreads();
  }
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&hdr->epoch);
  const unsigned tag = epoch % 0x3ffffffeu + 1u;

  // ---- stage the raw words this tile can touch (tile +- kh rows +- 1 word): one round of
  //      independent coalesced loads instead of (2kh+1) dependent ones per word ----------------
  const uint32_t* win = raw;
  int win0 = 0;
  if (!pool_hin) {
    const int lo = max(0, tile * kCompactTile - kh * Wd - 1);
    const int hi = min(nwords, (tile + 1) * kCompactTile + kh * Wd + 1);
    if (hi - lo <= kCompactWin) {
      for (int i = tid; i < hi - lo; i += kCompactThreads) s_win[i] = __ldg(raw + lo + i);
      win = s_win;
      win0 = lo;
      __syncthreads();
    }
  }

  // ---- dilated words (consecutive words per thread keep the index order) ------------------
  unsigned d[kCompactWPT];
  unsigned twin[kCompactWPT];                                // tiles this thread stamped first
  int tt0[kCompactWPT];
  int cnt = 0;
  const int w0 = tile * kCompactTile + tid * kCompactWPT;    // 32-bit index math: nwords < 2^31
#pragma unroll
  for (int i = 0; i < kCompactWPT; ++i) {
    const int w = w0 + i;
    d[i] = 0;
    twin[i] = 0;
    tt0[i] = 0;
    if (w < nwords) {
      const int j = w % Wd;
      const int r = w / Wd;
      const int y = r % H;
      d[i] = pool_hin ? pooled_word(raw, r / H, y, j, pool_hin, pool_wdin, W)
                      : dilated_word(win, win0, r, y, j, H, W, Wd, kh, kw);
      if (dil_bits) dil_bits[w] = d[i];
      if (HINTS && d[i]) prefetch_hinted_rows(hints, d[i], r / H, y, j);
    }
    if (tile_ws) {
      // dirty 8 x 16 output tiles for the tiled contraction (conv_tile.cuh): a word covers four
      // tiles of one tile row; the first word to stamp a tile with this launch's tag appends it
      // (unordered list, every tile once).  The four exchanges of a word are independent; their
      // results are only needed after the index expansion, where the append position is reserved
      // once per warp.
      uint32_t* stamp = reinterpret_cast<uint32_t*>(tile_ws) + 4;
      int t0 = 0;
      unsigned win = 0;
      if (w < nwords && d[i]) {
        const int j = w % Wd, r = w / Wd;
        t0 = ((r / H) * tile_ty + (r % H) / 16) * tile_xp + 4 * j;
        unsigned old[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          old[q] = ((d[i] >> (8 * q)) & 0xffu) ? atomicExch(stamp + t0 + q, tag) : tag;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (old[q] != tag) win |= 1u << q;
      }
      twin[i] = win;
      tt0[i] = t0;
    }
    cnt += __popc(d[i]);
  }

  if (no_list) {
    // tiles only (the consumer walks the tile list and masks rows with dil_bits): no ordered index
    // list, hence no scan across tiles -- the change count is a plain sum (hdr->reserved)
    int tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0 && tot) atomicAdd(&hdr->reserved, (unsigned)tot);
  } else {
  // ---- block scan of popcounts -----------------------------------------------------------
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = lane < NW ? s_warp[lane] : 0;
    int inc2 = v;
#pragma unroll
    for (int o = 1; o < NW; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += n;
    }
    if (lane < NW) s_warp[lane] = inc2 - v;                  // exclusive warp offsets
    const int total = __shfl_sync(0xffffffffu, inc2, NW - 1);

    // ---- chained scan across tiles (decoupled look-back), warp 0 ----------------------------
    int exclusive = 0;
    if (tile > 0) {
      if (lane == 0) tstate[tile] = pack_state(tag, 1u, (unsigned)total);    // aggregate
      int look = tile - 1;
      while (true) {
        const int t = look - lane;
        unsigned long long st = 0;
        if (t >= 0) {
          do { st = tstate[t]; } while ((unsigned)(st >> 34) != tag);
        }
        const unsigned status = t >= 0 ? (unsigned)((st >> 32) & 3u) : 2u;   // before tile 0: prefix 0
        const int val = t >= 0 ? (int)(unsigned)(st & 0xffffffffu) : 0;
        const unsigned pm = __ballot_sync(0xffffffffu, status == 2u);
        const int first = pm ? __ffs(pm) - 1 : 32;           // nearest tile with a full prefix
        int part = lane <= first ? val : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        exclusive += part;
        if (pm) break;
        look -= 32;
      }
    }
    if (lane == 0) {
      tstate[tile] = pack_state(tag, 2u, (unsigned)(exclusive + total));     // inclusive prefix
      s_base = exclusive;
      if (tile == ntiles - 1) *count = exclusive + total;
    }
  }
  __syncthreads();

  // ---- expand set bits to ascending pixel indices ------------------------------------------
  int o = s_base + s_warp[wid] + (incl - cnt);
  if (coop && kCompactWPT == 1 && !dil_map) {
    // warp-cooperative expansion: the warp walks its non-empty words; lane i writes bit i of the
    // current word at its rank -> the 32 stores of a dense word are one coalesced 128-byte line
    // (per-thread expansion writes 32 lines per instruction when a block of pixels changed)
    const int w = w0;
    const bool have = w < nwords && d[0] != 0u;
    const int jj = have ? w % Wd : 0, rr = have ? w / Wd : 0;
    const int mypix0 = rr * W + jj * 32;
    unsigned todo = __ballot_sync(0xffffffffu, have);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned dd = __shfl_sync(0xffffffffu, d[0], src);
      const int oo = __shfl_sync(0xffffffffu, o, src);
      const int pp = __shfl_sync(0xffffffffu, mypix0, src);
      if ((dd >> lane) & 1u) idx[oo + __popc(dd & ((1u << lane) - 1u))] = pp + lane;
    }
  } else {
#pragma unroll
  for (int i = 0; i < kCompactWPT; ++i) {
    const int w = w0 + i;
    if (w >= nwords) break;
    const int j = w % Wd;
    const int r = w / Wd;                                    // r = b*H + y
    const int pix0 = r * W + j * 32;
    unsigned dd = d[i];
    while (dd) {
      const int bit = __ffs(dd) - 1;
      idx[o++] = pix0 + bit;
      dd &= dd - 1;
    }
    if (dil_map) {
      const int n = min(32, W - j * 32);
      int8_t* m = dil_map + pix0;
      for (int q = 0; q < n; ++q) m[q] = (int8_t)((d[i] >> q) & 1u);
    }
  }
  }

  }   // !no_list

  // ---- append the tiles this warp stamped first ------------------------------------------------
  if (tile_ws) {
#pragma unroll
    for (int i = 0; i < kCompactWPT; ++i) {
      const unsigned win = twin[i];
      const int t0 = tt0[i];
      const int mine = __popc(win);
      int incl_t = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl_t, o);
        if (lane >= o) incl_t += v;
      }
      const int total_t = __shfl_sync(0xffffffffu, incl_t, 31);
      int base_t = 0;
      if (total_t) {
        if (lane == 31) base_t = atomicAdd(tile_ws, total_t);
        base_t = __shfl_sync(0xffffffffu, base_t, 31);
        int pos = 4 + B * tile_ty * tile_xp + base_t + incl_t - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if ((win >> q) & 1u) tile_ws[pos++] = t0 + q;
      }
    }
    __syncthreads();                            // all appends of this block precede its done-increment
  }

  // ---- leave the workspace clean for the next launch ---------------------------------------
  __shared__ unsigned s_last;
  if (tid == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&hdr->done, 1u);
    s_last = prev == (unsigned)ntiles - 1u;
    if (s_last) {
      hdr->done = 0;
      hdr->ticket = 0;
      if (no_list) {                            // every block added its popcounts before its done-increment
        *count = (int32_t)*reinterpret_cast<volatile unsigned*>(&hdr->reserved);
        hdr->reserved = 0;
      }
      if (tile_ws) {                            // every block appended before its done-increment
        tile_ws[1] = *reinterpret_cast<volatile int32_t*>(tile_ws);
        tile_ws[0] = 0;
      }
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(&hdr->epoch) = epoch + 1u;
    }
  }
  // The last tile to finish (every tile has read its raw window by then) may zero the raw
  // bitmap, so next frame's candidate detection can OR bits into it without a memset launch.
  if (clear_bits) {
    __syncthreads();
    if (s_last)
      for (int i = tid; i < nwords; i += kCompactThreads) clear_bits[i] = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// block_dilate_compact -- dilation + ordered compaction of a SMALL bitmap (<= kCompactWin words, e.g. one
// 368 x 368 map or eight 46 x 46 maps) by ONE thread block, called by the last block of the kernel that
// produced the raw bitmap (detect_sparse_vec_kernel<.., FUSE>): a layer on a small map then needs two
// launches per frame (detect + compact, contraction) instead of three -- such layers are bound by the
// dependent-launch chain, not by bytes (pose network: 92 layers of <= 4416 bitmap words at batch 1).
// Same results as dilate_compact_kernel: dilated bitmap, ascending index list, count; optional
// zeroing of the raw bitmap.  All threads of the block must call it; `s_win` holds kCompactWin words.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_dilate_compact(uint32_t* __restrict__ raw, uint32_t* __restrict__ dil_bits,
                                                     int32_t* __restrict__ idx, int32_t* __restrict__ count,
                                                     uint32_t* s_win, int* s_warp, int H, int W, int Wd,
                                                     int kh, int kw, int nwords, bool clear) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarp = nthr >> 5;
  for (int i = tid; i < nwords; i += nthr) s_win[i] = __ldcg(raw + i);     // (written by other blocks: L2)
  __syncthreads();
  const int wpt = (nwords + nthr - 1) / nthr;                              // consecutive words per thread
  const int w0 = tid * wpt, w1 = min(nwords, w0 + wpt);
  int cnt = 0;
  for (int w = w0; w < w1; ++w) {
    const int j = w % Wd, r = w / Wd;
    const unsigned d = dilated_word(s_win, 0, r, r % H, j, H, W, Wd, kh, kw);
    if (dil_bits) dil_bits[w] = d;
    cnt += __popc(d);
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int woff = 0, total = 0;
  for (int q = 0; q < nwarp; ++q) {
    const int v = s_warp[q];
    if (q < wid) woff += v;
    total += v;
  }
  int o = woff + incl - cnt;
  for (int w = w0; w < w1; ++w) {
    const int j = w % Wd, r = w / Wd;
    unsigned d = dilated_word(s_win, 0, r, r % H, j, H, W, Wd, kh, kw);
    const int pix0 = r * W + j * 32;
    while (d) {
      idx[o++] = pix0 + __ffs(d) - 1;
      d &= d - 1;
    }
  }
  if (tid == 0) *count = total;
  if (clear)
    for (int i = tid; i < nwords; i += nthr) raw[i] = 0u;
}

// ------------------------------------------------------------------------------------------------
// dilate_tiles_kernel -- dilation + dirty-tile list + change count, for consumers that walk tiles
// (cb_dilate_tiles; no ordered index list, hence no scan across blocks).
//
// One block per (image, tile row) = 16 map rows x Wd bitmap words: the raw rows it can touch
// (16 + 2*kh, zero-padded by one word on each side) are staged in shared memory with one round of
// independent loads; a thread computes a dilated word from shared memory, stores it and ORs it into
// its column's accumulator; the column accumulators are the tile flags (a word covers four 8-pixel
// tiles), so every tile has exactly ONE owner thread -- no stamps, no exchange atomics.  ONE 64-bit
// atomicAdd per block carries (blocks done | tiles appended | changed pixels): its return value is
// both the block's append position and the "am I last" test, and the last block publishes the totals
// from it without reading anything back.  Two dependent global round trips per block (stage, atomic)
// instead of five in dilate_compact_kernel's tile mode.
// Clearing the raw bitmap (candidate-path layers): a block zeroes the rows only it reads right after
// staging them; the rows shared with the neighbouring tile rows are left to the last block.
//   sync word layout: [63:51] blocks done | [50:31] tiles | [30:0] changed pixels   (0 at rest)
// ------------------------------------------------------------------------------------------------
constexpr int kDtThreads = 256;
constexpr int kDtRows = 16;                                   // = TL_H of conv_tile.cuh

__host__ __device__ inline size_t dilate_tiles_smem(int Wd, int kh) {
  return ((size_t)(kDtRows + 2 * kh) * (Wd + 2) + Wd) * sizeof(uint32_t);
}
// limits of the packed counter; beyond them cb_dilate_tiles keeps dilate_compact_kernel
__host__ __device__ inline bool dilate_tiles_ok(int B, int H, int W, int kh, int tile_ty, int tile_xp) {
  const int Wd = (W + 31) / 32;
  return kh <= kDtRows / 2 && (long long)B * tile_ty < 8192 && (long long)B * tile_ty * tile_xp < (1ll << 20) &&
         dilate_tiles_smem(Wd, kh) <= 48 * 1024;
}

__global__ void __launch_bounds__(kDtThreads)
dilate_tiles_kernel(const uint32_t* __restrict__ raw, uint32_t* __restrict__ dil_bits,
                    int32_t* __restrict__ count, unsigned long long* __restrict__ sync, int B, int H,
                    int W, int Wd, int kh, int kw, uint32_t* __restrict__ clear_bits,
                    int32_t* __restrict__ tile_ws, int tile_ty, int tile_xp) {
  pdl_prologue();
  extern __shared__ uint32_t dt_smem[];
  const int WS = Wd + 2, NR = kDtRows + 2 * kh;
  uint32_t* s_raw = dt_smem;                                  // [NR][WS], column 0 and WS-1 are zero
  uint32_t* s_col = dt_smem + NR * WS;                        // [Wd] OR of the block's dilated words
  __shared__ int s_warp[kDtThreads / 32], s_cnt[kDtThreads / 32];
  __shared__ unsigned long long s_old;
  __shared__ int s_ntl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.x / tile_ty, ty = blockIdx.x - b * tile_ty;
  const int y0 = ty * kDtRows;
  const uint32_t* rimg = raw + (long long)b * H * Wd;

  // ---- stage rows y0-kh .. y0+15+kh (zero outside the image) ---------------------------------
  for (int i = tid; i < NR * WS; i += kDtThreads) {
    const int r = i / WS, c = i - r * WS;
    const int y = y0 - kh + r, j = c - 1;
    s_raw[i] = (y >= 0 && y < H && j >= 0 && j < Wd) ? __ldg(rimg + (long long)y * Wd + j) : 0u;
  }
  for (int i = tid; i < Wd; i += kDtThreads) s_col[i] = 0u;
  __syncthreads();
  if (clear_bits) {                                           // rows nobody else reads: mine to clear
    uint32_t* cimg = clear_bits + (long long)b * H * Wd;
    const int lo = ty == 0 ? 0 : kh, hi = ty == tile_ty - 1 ? kDtRows : kDtRows - kh;
    for (int i = tid; i < (hi - lo) * Wd; i += kDtThreads) {
      const int y = y0 + lo + i / Wd;
      if (y < H) cimg[(long long)y * Wd + i % Wd] = 0u;
    }
  }

  // ---- dilated words ---------------------------------------------------------------------------
  int cnt = 0;
  uint32_t* dimg = dil_bits + (long long)b * H * Wd;
  for (int i = tid; i < kDtRows * Wd; i += kDtThreads) {
    const int r = i / Wd, j = i - r * Wd;
    const int y = y0 + r;
    if (y >= H) break;
    unsigned vp = 0, vc = 0, vn = 0;
    const uint32_t* p = s_raw + r * WS + j;                   // row y-kh, column j-1
    for (int q = 0; q <= 2 * kh; ++q, p += WS) {
      vp |= p[0];
      vc |= p[1];
      vn |= p[2];
    }
    unsigned d = vc;
    for (int dx = 1; dx <= kw; ++dx) {
      d |= (vc << dx) | (vp >> (32 - dx));
      d |= (vc >> dx) | (vn << (32 - dx));
    }
    if (j == Wd - 1 && (W & 31)) d &= (1u << (W & 31)) - 1u;
    dimg[(long long)y * Wd + j] = d;
    if (d) {
      atomicOr(&s_col[j], d);
      cnt += __popc(d);
    }
  }
  __syncthreads();

  // ---- tiles of this block: thread j owns the four tiles of column word j --------------------------
  // (maps wider than 32 * kDtThreads pixels take several rounds of columns)
  unsigned flags = 0;
  int mine = 0;
  for (int c0 = 0; c0 < Wd; c0 += kDtThreads) {
    const int j = c0 + tid;
    unsigned f = 0;
    if (j < Wd) {
      const unsigned v = s_col[j];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((v >> (8 * q)) & 0xffu) f |= 1u << q;
    }
    if (c0 == 0) flags = f;                                   // first round is kept in registers
    mine += __popc(f);
  }
  // block totals: tiles (prefix needed) and changed pixels
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  int csum = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
  if (lane == 31) s_warp[wid] = incl;
  if (lane == 0) s_cnt[wid] = csum;
  __syncthreads();
  int woff = 0, ntl_blk = 0, cnt_blk = 0;
#pragma unroll
  for (int q = 0; q < kDtThreads / 32; ++q) {
    if (q < wid) woff += s_warp[q];
    ntl_blk += s_warp[q];
    cnt_blk += s_cnt[q];
  }
  const unsigned nblocks = (unsigned)B * tile_ty;
  if (tid == 0) {
    __threadfence();                                          // (clears and dil_bits precede the count)
    s_old = atomicAdd(sync, (1ull << 51) | ((unsigned long long)ntl_blk << 31) | (unsigned long long)cnt_blk);
  }
  __syncthreads();
  const unsigned long long old = s_old;
  const int base = (int)((old >> 31) & 0xFFFFFull);
  const bool last = (unsigned)(old >> 51) == nblocks - 1u;
  {
    int32_t* list = tile_ws + 4 + B * tile_ty * tile_xp;
    int pos = base + woff + incl - mine;
    const int t0 = (b * tile_ty + ty) * tile_xp;
    for (int c0 = 0; c0 < Wd; c0 += kDtThreads) {
      const int j = c0 + tid;
      unsigned f = flags;
      if (c0 > 0) {
        f = 0;
        if (j < Wd) {
          const unsigned v = s_col[j];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if ((v >> (8 * q)) & 0xffu) f |= 1u << q;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((f >> q) & 1u) list[pos++] = t0 + 4 * j + q;
    }
  }
  if (last) {
    if (tid == 0) {
      tile_ws[1] = base + ntl_blk;
      *count = (int32_t)(old & 0x7FFFFFFFull) + cnt_blk;
      *sync = 0ull;
    }
    if (clear_bits && kh > 0) {                               // the rows two tile rows share
      const int slots = (int)nblocks * 2 * kh;                // (image, tile row) x (top kh rows, bottom kh rows)
      for (int i = tid; i < slots * Wd; i += kDtThreads) {    // (32-bit: < the bitmap's word count)
        const int row = i / Wd, j = i - row * Wd;
        const int blk = row / (2 * kh), q = row - blk * 2 * kh;
        const int bb = blk / tile_ty, tt = blk - bb * tile_ty;
        if ((q < kh && tt == 0) || (q >= kh && tt == tile_ty - 1)) continue;   // cleared by their owner
        const int y = tt * kDtRows + (q < kh ? q : kDtRows - 2 * kh + q);
        if (y < H) clear_bits[((long long)bb * H + y) * Wd + j] = 0u;
      }
    }
  }
}

__global__ void map_to_bits_kernel(const int8_t* __restrict__ map, uint32_t* __restrict__ bits,
                                   int H, int W, int Wd, long long nwords) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nwords) return;
  const int j = (int)(w % Wd);
  const long long r = w / Wd;
  const int x = j * 32 + lane;
  const bool f = x < W && map[r * W + x] != 0;
  const unsigned word = __ballot_sync(0xffffffffu, f);
  if (lane == 0) bits[w] = word;
}

}  // namespace cb
