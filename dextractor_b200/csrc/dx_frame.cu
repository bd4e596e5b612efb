// dx_frame.cu -- text framing on the device.
//
// Replaces the fgets()/strlen() line loops of the reference (Read_Lines, QV.c:751-798; the
// sequence-line loop of dexta.c:161-183) by an ordered index of "interesting" byte positions:
//   DX_PRED_NEWLINE    every '\n'
//   DX_PRED_FASTA_HDR  every '>' that starts a line
//   DX_PRED_QVCAND     every offset whose next 12 bytes look like the beg/end/qv fields of a
//                      .dexqv entry header (dexqv.c:137-139) -- candidates only, verified later
// One pass over the text (per-tile slot rows, scan, gather); the exact form -- per-tile counts, an
// exclusive scan of the counts, an ordered write in a second pass -- remains as the fallback for
// inputs with more than 16 hits in a 16 KB tile.

#include <stdlib.h>
#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kTileThreads = 256;
constexpr int kTileChunks  = 4;                                  // 16-byte chunks per thread
constexpr int kTileBytes   = kTileThreads * kTileChunks * 16;    // 16 KB

// plausibility window for a .dexqv entry header (see dx_api.cpp: a true entry outside this
// window is still found by the verified chain walk, only slower)
constexpr uint32_t kCandMaxBeg  = 1u << 27;
constexpr uint32_t kCandMaxLen  = 1u << 20;
constexpr uint32_t kCandMaxQv   = 1u << 16;

__device__ __forceinline__ uint32_t load_le32(const uint8_t *p)
{ return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }

// 16-bit hit mask of the chunk starting at byte `at` (16-byte aligned, at < n)
template <int PRED>
__device__ __forceinline__ uint32_t chunk_hits(const uint8_t *buf, size_t n, size_t first, size_t at, uint4 v);

template <int PRED>
__device__ __forceinline__ uint32_t chunk_hits(const uint8_t *buf, size_t n, size_t first, size_t at)
{ return chunk_hits<PRED>(buf,n,first,at,dx_ldg16(buf + at));   // the buffer is padded to a 16-byte multiple
}

template <int PRED>
__device__ __forceinline__ uint32_t chunk_hits(const uint8_t *buf, size_t n, size_t first, size_t at, uint4 v)
{ uint32_t hits;
  if (PRED == DX_PRED_NEWLINE)
    { // newlines are rare: one exact "is any byte a newline" test for the whole chunk first
      const uint32_t k = 0x0a0a0a0au;
      const uint32_t a = v.x ^ k, b = v.y ^ k, c = v.z ^ k, d = v.w ^ k;
      const uint32_t z = ((a - 0x01010101u) & ~a) | ((b - 0x01010101u) & ~b) |
                         ((c - 0x01010101u) & ~c) | ((d - 0x01010101u) & ~d);
      if ((z & 0x80808080u) == 0) return 0;
      hits = dx_eq_mask16(v,'\n');
    }
  else if (PRED == DX_PRED_FASTA_HDR)
    { uint32_t gt = dx_eq_mask16(v,'>');
      if (gt == 0) return 0;
      uint32_t nl = dx_eq_mask16(v,'\n') << 1;
      if (at == 0 || buf[at-1] == '\n') nl |= 1u;
      hits = gt & nl;
    }
  else if (PRED == DX_PRED_ARCAND)
    { // .dexar entry: int32 beg, int32 end, 4 x uint16 SNR (<= 9999 each)  (dexar.c:202-204)
      uint4 w = (at + 16 < n) ? dx_ldg16(buf + at + 16) : make_uint4(~0u,~0u,~0u,~0u);
      const uint32_t k = 0xf8f8f8f8u;
      uint4 vm = make_uint4(v.x & k,v.y & k,v.z & k,v.w & k);
      uint4 wm = make_uint4(w.x & k,w.y & k,w.z & k,w.w & k);
      uint32_t z = dx_eq_mask16(vm,0) | (dx_eq_mask16(wm,0) << 16);    // bytes <= 7
      uint32_t m = (z >> 3) & (z >> 7) & 0xffffu;
      hits = 0;
      while (m)
        { int i = __ffs(m) - 1;
          m &= m - 1;
          size_t p = at + i;
          if (p + 16 > n) break;
          uint32_t beg = load_le32(buf+p), end = load_le32(buf+p+4);
          uint32_t s01 = load_le32(buf+p+8), s23 = load_le32(buf+p+12);
          if (beg < kCandMaxBeg && end >= beg && end - beg <= kCandMaxLen &&
              (s01 & 0xffffu) <= 9999u && (s01 >> 16) <= 9999u &&
              (s23 & 0xffffu) <= 9999u && (s23 >> 16) <= 9999u)
            hits |= 1u << i;
        }
    }
  else
    { // cheap filter: bytes +10 and +11 (top half of qv) must be zero; then the full test
      uint4 w = (at + 16 < n) ? dx_ldg16(buf + at + 16) : make_uint4(~0u,~0u,~0u,~0u);
      uint32_t z = dx_eq_mask16(v,0) | (dx_eq_mask16(w,0) << 16);
      uint32_t m = (z >> 10) & (z >> 11) & 0xffffu;
      hits = 0;
      while (m)
        { int i = __ffs(m) - 1;
          m &= m - 1;
          size_t p = at + i;
          if (p + 12 > n) break;
          uint32_t beg = load_le32(buf+p), end = load_le32(buf+p+4), qv = load_le32(buf+p+8);
          if (beg < kCandMaxBeg && end >= beg && end - beg <= kCandMaxLen && qv < kCandMaxQv)
            hits |= 1u << i;
        }
    }
  // clip to [first, n)
  int lo = (first > at) ? (int) min((size_t) 16, first - at) : 0;
  int hi = (n - at < 16) ? (int) (n - at) : 16;
  return hits & dx_range16(lo,hi);
}

template <int PRED>
__global__ void __launch_bounds__(kTileThreads)
k_pred_count(const uint8_t *buf, size_t n, size_t first, uint32_t *tile_count)
{ const size_t base = (size_t) blockIdx.x * kTileBytes;
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kTileChunks; j++)
    { size_t at = base + ((size_t) j * kTileThreads + threadIdx.x) * 16;
      if (at < n)
        cnt += __popc(chunk_hits<PRED>(buf,n,first,at));
    }
  __shared__ uint32_t part[kTileThreads/32];
  cnt = dx_warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0)
    { uint32_t s = 0;
      for (int w = 0; w < kTileThreads/32; w++) s += part[w];
      tile_count[blockIdx.x] = s;
    }
}

// exclusive scan of ntiles 32-bit counts into 64-bit offsets; one CTA, 32 consecutive values per
// thread per round (vector loads and stores when the arrays are 16-byte aligned; a round costs a
// few microseconds of barrier and memory latency whatever it holds).  total at prefix[ntiles].
constexpr int kScanItems = 32;

__global__ void __launch_bounds__(1024)
k_tile_scan(const uint32_t *count, int64_t ntiles, int64_t *prefix)
{ __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = ((reinterpret_cast<uintptr_t>(count) | reinterpret_cast<uintptr_t>(prefix)) & 15) == 0;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < ntiles; b += 1024*kScanItems)
    { const int64_t i0 = b + (int64_t) threadIdx.x * kScanItems;
      uint32_t c[kScanItems];
      uint64_t v = 0;
      if (vec && i0 + kScanItems <= ntiles)
        {
#pragma unroll
          for (int k = 0; k < kScanItems; k += 4)
            { const uint4 q = *reinterpret_cast<const uint4 *>(count + i0 + k);
              c[k] = q.x; c[k+1] = q.y; c[k+2] = q.z; c[k+3] = q.w;
            }
        }
      else
        {
#pragma unroll
          for (int k = 0; k < kScanItems; k++) c[k] = (i0 + k < ntiles) ? count[i0 + k] : 0u;
        }
#pragma unroll
      for (int k = 0; k < kScanItems; k++) v += c[k];
      uint64_t inc = dx_warp_incl_sum64(v,lane);
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      if (warp == 0)
        { uint64_t w = wsum[lane];
          uint64_t wi = dx_warp_incl_sum64(w,lane);
          wsum[lane] = wi - w;                     // exclusive over warps
        }
      __syncthreads();
      uint64_t excl = carry + wsum[warp] + inc - v;
      if (vec && i0 + kScanItems <= ntiles)
        {
#pragma unroll
          for (int k = 0; k < kScanItems; k += 2)
            { const uint64_t x0 = excl, x1 = excl + c[k];
              *reinterpret_cast<ulonglong2 *>(prefix + i0 + k) = make_ulonglong2(x0,x1);
              excl = x1 + c[k+1];
            }
        }
      else
        {
#pragma unroll
          for (int k = 0; k < kScanItems; k++)
            { if (i0 + k < ntiles) prefix[i0 + k] = (int64_t) excl;
              excl += c[k];
            }
        }
      __syncthreads();
      if (threadIdx.x == 1023) carry = excl;
      __syncthreads();
    }
  if (threadIdx.x == 0) prefix[ntiles] = (int64_t) carry;
}

// ... the same scan over several CTAs, for more than a few thousand values: a tile of 8192 values per
// CTA (tiles handed out by a ticket, so a tile's predecessors are always running or done), every
// tile publishes its total as soon as it has it, and a tile's base is the sum of its predecessors'
// totals, gathered 32 at a time by its first warp.  No tile waits for anything but that first phase
// of earlier tiles.  state: [0] ticket, [1 + t] total of tile t | 1 << 63 once published (zeroed).
constexpr int kChainItems = 8;
constexpr int kChainTile  = 1024 * kChainItems;    // with the usual 1024 threads (tests run it with 32)

__global__ void __launch_bounds__(1024)
k_chain_scan(const uint32_t *count, int64_t n, int64_t *prefix, unsigned long long *state)
{ __shared__ uint64_t wsum[32];
  __shared__ uint64_t s_base, s_total;
  __shared__ unsigned long long s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long kReady = 1ull << 63;
  if (threadIdx.x == 0) s_tile = atomicAdd(&state[0],1ull);
  __syncthreads();
  const int64_t t = (int64_t) s_tile;
  const int64_t tile = (int64_t) blockDim.x * kChainItems;
  const int nwarp = (int) (blockDim.x >> 5);
  const int64_t i0 = t * tile + (int64_t) threadIdx.x * kChainItems;
  uint32_t c[kChainItems];
  uint64_t v = 0;
#pragma unroll
  for (int k = 0; k < kChainItems; k++) { c[k] = (i0 + k < n) ? count[i0 + k] : 0u; v += c[k]; }
  const uint64_t inc = dx_warp_incl_sum64(v,lane);
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0)
    { const uint64_t w = (lane < nwarp) ? wsum[lane] : 0ull;
      const uint64_t wi = dx_warp_incl_sum64(w,lane);
      wsum[lane] = wi - w;                          // exclusive over warps
      if (lane == 31)
        { s_total = wi;
          *reinterpret_cast<volatile unsigned long long *>(&state[1 + t]) = (unsigned long long) wi | kReady;
        }
    }
  __syncthreads();
  if (warp == 0)                                    // base = totals of tiles 0 .. t-1
    { uint64_t base = 0;
      for (int64_t p0 = 0; p0 < t; p0 += 32)
        { const int64_t p = p0 + lane;
          unsigned long long f = kReady;
          if (p < t)
            do f = *reinterpret_cast<volatile unsigned long long *>(&state[1 + p]); while (!(f & kReady));
          base += (p < t) ? (uint64_t) (f & ~kReady) : 0ull;
        }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) base += __shfl_xor_sync(DX_FULL,base,d);
      if (lane == 0) s_base = base;
    }
  __syncthreads();
  uint64_t excl = s_base + wsum[warp] + inc - v;
#pragma unroll
  for (int k = 0; k < kChainItems; k++)
    { if (i0 + k < n) prefix[i0 + k] = (int64_t) excl;
      excl += c[k];
    }
  if ((t + 1) * tile >= n && threadIdx.x == 0) prefix[n] = (int64_t) (s_base + s_total);
}

// exclusive scan of n 32-bit values into 64-bit offsets, total at prefix[n]
static int launch_scan(dx_ctx *ctx, const uint32_t *d_in, int64_t n, int64_t *d_prefix, const char *what)
{ const bool small_tiles = (ctx->route[DXR_CHAIN_SCAN] != 0);       // tests: tiles of 256 values
  if (n == 0 || (n <= 2*kChainTile && !small_tiles))
    { DX_PROF_BEGIN(ctx); k_tile_scan<<<1,1024,0,ctx->stream>>>(d_in,n,d_prefix);
      DX_LAUNCHED(ctx,what);
      return DX_OK;
    }
  const int threads = small_tiles ? 32 : 1024;
  const int64_t tile = (int64_t) threads * kChainItems;
  const int64_t ntile = (n + tile - 1) / tile;
  unsigned long long *d_state = (unsigned long long *) dx_arena_get(ctx,(size_t) (ntile + 1)*8);
  if (d_state == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_state,0,(size_t) (ntile + 1)*8,ctx->stream));
  DX_PROF_BEGIN(ctx); k_chain_scan<<<(unsigned) ntile,threads,0,ctx->stream>>>(d_in,n,d_prefix,d_state);
  DX_LAUNCHED(ctx,what);
  return DX_OK;
}

template <int PRED>
__global__ void __launch_bounds__(kTileThreads)
k_pred_write(const uint8_t *buf, size_t n, size_t first, const int64_t *tile_prefix, int64_t *pos)
{ __shared__ uint32_t wsum[kTileThreads/32];
  __shared__ uint32_t running;
  const size_t base = (size_t) blockIdx.x * kTileBytes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  int64_t *dst = pos + tile_prefix[blockIdx.x];
  for (int j = 0; j < kTileChunks; j++)
    { size_t at = base + ((size_t) j * kTileThreads + threadIdx.x) * 16;
      uint32_t hits = (at < n) ? chunk_hits<PRED>(buf,n,first,at) : 0;
      uint32_t c = __popc(hits);
      uint32_t inc = dx_warp_incl_sum(c,lane);
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      uint32_t before = running;
      for (int w = 0; w < warp; w++) before += wsum[w];
      uint32_t r = before + inc - c;
      while (hits)
        { int i = __ffs(hits) - 1;
          hits &= hits - 1;
          dst[r++] = (int64_t) (at + i);
        }
      __syncthreads();
      if (threadIdx.x == kTileThreads-1) running = before + inc;
      __syncthreads();
    }
}

// ---- .quiva entry table ------------------------------------------------------------------------

__device__ __forceinline__ bool take_digits(const uint8_t *t, int64_t &p, int64_t end, int32_t &val)
{ int64_t s = p;
  uint32_t v = 0;
  while (p < end && t[p] >= '0' && t[p] <= '9' && p - s < 9)
    v = v*10 + (t[p++] - '0');
  if (p == s || (p < end && t[p] >= '0' && t[p] <= '9')) return false;   // none, or > 9 digits
  val = (int32_t) v;
  return true;
}

// Canonical "<well>/<beg>_<end> RQ=0.<qv>" after the first '/' of the header (QV.c:958-968).
// Anything else (signs, blanks, > 9 digits, missing RQ) is left to the host's sscanf.
__device__ bool parse_quiva_header(const uint8_t *t, int64_t p, int64_t end,
                                   int32_t &well, int32_t &beg, int32_t &en, int32_t &qv)
{ p += 1;
  while (p < end && t[p] != '/') p++;
  if (p >= end) return false;
  p++;
  if (!take_digits(t,p,end,well) || p >= end || t[p] != '/') return false;
  p++;
  if (!take_digits(t,p,end,beg) || p >= end || t[p] != '_') return false;
  p++;
  if (!take_digits(t,p,end,en)) return false;
  if (p + 6 > end || t[p] != ' ' || t[p+1] != 'R' || t[p+2] != 'Q' || t[p+3] != '=' ||
      t[p+4] != '0' || t[p+5] != '.') return false;
  p += 6;
  return take_digits(t,p,end,qv);
}

__global__ void k_qv_entries(const uint8_t *text, const int64_t *nl, int64_t nent, QvEntries ent,
                             unsigned long long *err /*[0] first error, [1] total positions, [2] headers
                                                       the host has to parse, [3] offset of the last newline*/)
{ int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  { // total positions of the shard (totChar, QV.c:1005): one atomic per warp
    uint32_t mylen = 0;
    if (e < nent)
      { int64_t l = nl[6*e+1] - nl[6*e] - 1;
        mylen = (l > 0 && l < 0x7ffffff0) ? (uint32_t) l : 0;
      }
    unsigned long long s = dx_warp_incl_sum64(mylen,threadIdx.x & 31);
    if ((threadIdx.x & 31) == 31 && s) atomicAdd(err+1,s);
  }
  if (e >= nent) return;
  if (e == nent-1) err[3] = (unsigned long long) nl[6*e+5];
  const int64_t h0 = (e == 0) ? 0 : nl[6*e-1] + 1;
  const int64_t h1 = nl[6*e];
  int64_t len = nl[6*e+1] - h1 - 1;
  unsigned long long bad = 0;
  if (h1 == h0 || text[h0] != '@')
    bad = ((unsigned long long) (6*e+1) << 8) | 1;                // header missing
  for (int k = 2; k <= 5 && !bad; k++)
    if (nl[6*e+k] - nl[6*e+k-1] - 1 != len)
      bad = ((unsigned long long) (6*e+k+1) << 8) | 5;            // lines differ in length
  if (len >= (1 << 24)) bad = ((unsigned long long) (6*e+2) << 8) | 6;   // beyond 32-bit bit offsets
  if (bad) { atomicMin(err,bad); return; }
  ent.hdr[e]   = h0;
  ent.line0[e] = h1 + 1;
  ent.rlen[e]  = (int32_t) len;
  int32_t well = 0, beg = 0, en = 0, qv = 0;
  bool ok = parse_quiva_header(text,h0,h1,well,beg,en,qv);
  ent.well[e] = well; ent.beg[e] = beg; ent.end[e] = en; ent.qv[e] = qv;
  ent.flag[e] = ok ? 0 : 1;
  if (!ok) atomicAdd(err+2,1ull);
}

}  // namespace

// Small results for the host WITHOUT the copy engine: a kernel stores them into pinned host memory
// (device-accessible under UVA).  A cudaMemcpy D2H of a few bytes queues on the device-to-host copy
// engine BEHIND whatever large copy is in flight there -- in the pipelined dx_undexqv_host that
// serialised every window's planning with the text copy of the window before it.
__global__ void k_fetch(uint8_t *dst, const uint8_t *src, size_t n)
{ const size_t i0 = (size_t) blockIdx.x * blockDim.x + threadIdx.x, step = (size_t) gridDim.x * blockDim.x;
  if ((((uintptr_t) dst | (uintptr_t) src) & 15) == 0)
    { const size_t nv = n >> 4;
      for (size_t i = i0; i < nv; i += step)
        reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
      for (size_t i = (nv << 4) + i0; i < n; i += step) dst[i] = src[i];
    }
  else
    for (size_t i = i0; i < n; i += step) dst[i] = src[i];
}

int dxk_fetch(dx_ctx *ctx, void *h_pinned, const void *d_src, size_t bytes)
{ if (bytes == 0) return DX_OK;
  size_t blocks = (bytes/16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 1024) blocks = 1024;
  k_fetch<<<(unsigned) blocks,256,0,ctx->stream>>>((uint8_t *) h_pinned,(const uint8_t *) d_src,bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return dx_cuda_fail(ctx,e,"k_fetch");
  return DX_OK;
}

namespace {

template <int PRED>
int index_positions_exact(dx_ctx *ctx, const uint8_t *buf, size_t n, size_t first,
                          int64_t **d_pos, int64_t *count)
{ const int64_t ntiles = (int64_t) ((n + kTileBytes - 1) / kTileBytes);
  *d_pos = NULL; *count = 0;
  if (ntiles == 0) return DX_OK;
  uint32_t *d_cnt = (uint32_t *) dx_arena_get(ctx,(size_t) ntiles*4);
  int64_t  *d_pre = (int64_t *)  dx_arena_get(ctx,(size_t) (ntiles+1)*8);
  if (d_cnt == NULL || d_pre == NULL) return DX_E_NOMEM;
  DX_PROF_BEGIN(ctx); k_pred_count<PRED><<<(unsigned) ntiles,kTileThreads,0,ctx->stream>>>(buf,n,first,d_cnt);
  DX_LAUNCHED(ctx,"k_pred_count");
  { const int rc = launch_scan(ctx,d_cnt,ntiles,d_pre,"k_tile_scan");
    if (rc != DX_OK) return rc;
  }
  int64_t total = 0;
  DX_CUDA(ctx,cudaMemcpyAsync(&total,d_pre+ntiles,8,cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  *count = total;
  if (total == 0) return DX_OK;
  int64_t *pos = (int64_t *) dx_arena_get(ctx,(size_t) total*8);
  if (pos == NULL) return DX_E_NOMEM;
  DX_PROF_BEGIN(ctx); k_pred_write<PRED><<<(unsigned) ntiles,kTileThreads,0,ctx->stream>>>(buf,n,first,d_pre,pos);
  DX_LAUNCHED(ctx,"k_pred_write");
  *d_pos = pos;
  return DX_OK;
}

// ---- one pass over the text: per-tile slots ---------------------------------------------------------
// Hits are sparse (a newline every ~10 KB of .quiva, a header every ~10 KB of .fasta, an entry every
// ~17 KB of .dexqv): every 16 KB tile drops its (few) hit positions, unordered, into a fixed row of
// kSlot slots and reports its count; an exclusive scan of the counts and a gather that sorts each
// row give the ordered index.  The text is read once, by a kernel as simple as k_pred_count.  A tile
// with more than kSlot hits (dense input) sends the call to the exact two-pass kernels above.
constexpr int kSlot = 16;

template <int PRED>
__global__ void __launch_bounds__(kTileThreads)
k_pred_slots(const uint8_t *buf, size_t n, size_t first, uint32_t *tile_count, int64_t *slots, int32_t *overflow)
{ __shared__ uint32_t cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const size_t base = (size_t) blockIdx.x * kTileBytes;
  uint4 v[kTileChunks];
#pragma unroll
  for (int j = 0; j < kTileChunks; j++)
    { const size_t at = base + ((size_t) j * kTileThreads + threadIdx.x) * 16;
      v[j] = (at < n) ? dx_ldg16(buf + at) : make_uint4(0,0,0,0);
    }
#pragma unroll
  for (int j = 0; j < kTileChunks; j++)
    { const size_t at = base + ((size_t) j * kTileThreads + threadIdx.x) * 16;
      uint32_t hits = (at < n) ? chunk_hits<PRED>(buf,n,first,at,v[j]) : 0;
      if (hits)
        { uint32_t s = atomicAdd(&cnt,(uint32_t) __popc(hits));
          while (hits)
            { const int i = __ffs(hits) - 1;
              hits &= hits - 1;
              if (s < (uint32_t) kSlot) slots[(size_t) blockIdx.x * kSlot + s] = (int64_t) (at + i);
              s++;
            }
        }
    }
  __syncthreads();
  if (threadIdx.x == 0)
    { tile_count[blockIdx.x] = cnt;
      if (cnt > (uint32_t) kSlot) atomicExch(overflow,1);
    }
}

// ---- the same pass with TMA bulk copies (an experiment, route "index_bulk") ----------------------
// One elected thread per CTA moves whole 16 KB tiles global -> shared with cp.async.bulk and an
// mbarrier (UBLKCP in the SASS), two tiles in flight, every thread then reads its four chunks from
// shared memory.  This takes the LDG instructions and their address arithmetic off the 256 threads
// and hands the data movement to the copy engine of the SM -- what the round-1 verdict asked to try.
// Measured on the 2 GB text (profiles/r02_tma_experiment.txt): see there; the LDG form above already
// runs at the measured HBM peak, which is the most a different way of fetching the same bytes can do.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }

__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{ asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@!p bra WAIT_%=;\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int PRED>
__global__ void __launch_bounds__(kTileThreads)
k_pred_slots_bulk(const uint8_t *buf, size_t n, size_t first, int64_t ntiles, uint32_t *tile_count, int64_t *slots,
                  int32_t *overflow)
{ extern __shared__ __align__(128) uint8_t dx_bulk_smem[];           // 2 x 16 KB tiles
  __shared__ uint64_t bar[2];
  __shared__ uint32_t cnt;
  if (threadIdx.x == 0)
    { mbar_init(&bar[0],1); mbar_init(&bar[1],1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();
  auto tile_bytes = [&](int64_t t) -> uint32_t
    { const size_t base = (size_t) t * kTileBytes;
      const size_t left = n - base;
      return (uint32_t) (((left < (size_t) kTileBytes ? left : (size_t) kTileBytes) + 15) & ~(size_t) 15);
    };
  int64_t t = blockIdx.x;
  if (threadIdx.x == 0 && t < ntiles) bulk_load(dx_bulk_smem,buf + (size_t) t*kTileBytes,tile_bytes(t),&bar[0]);
  uint32_t phase[2] = { 0, 0 };
  for (int it = 0; t < ntiles; t += gridDim.x, it++)
    { const int b = it & 1;
      const int64_t tn = t + gridDim.x;
      if (threadIdx.x == 0)
        { cnt = 0;
          if (tn < ntiles)                                        // the other buffer was drained an iteration ago
            bulk_load(dx_bulk_smem + (size_t) (b ^ 1)*kTileBytes,buf + (size_t) tn*kTileBytes,tile_bytes(tn),&bar[b ^ 1]);
        }
      mbar_wait(&bar[b],phase[b]); phase[b] ^= 1u;
      __syncthreads();                                            // cnt = 0 is visible
      const size_t base = (size_t) t * kTileBytes;
      const uint4 *tile = reinterpret_cast<const uint4 *>(dx_bulk_smem + (size_t) b*kTileBytes);
#pragma unroll
      for (int j = 0; j < kTileChunks; j++)
        { const size_t at = base + ((size_t) j * kTileThreads + threadIdx.x) * 16;
          uint32_t hits = (at < n) ? chunk_hits<PRED>(buf,n,first,at,tile[j * kTileThreads + threadIdx.x]) : 0;
          if (hits)
            { uint32_t sidx = atomicAdd(&cnt,(uint32_t) __popc(hits));
              while (hits)
                { const int i = __ffs(hits) - 1;
                  hits &= hits - 1;
                  if (sidx < (uint32_t) kSlot) slots[(size_t) t * kSlot + sidx] = (int64_t) (at + i);
                  sidx++;
                }
            }
        }
      __syncthreads();                                            // everybody is done with buffer b and with cnt
      if (threadIdx.x == 0)
        { tile_count[t] = cnt;
          if (cnt > (uint32_t) kSlot) atomicExch(overflow,1);
        }
    }
}

__global__ void k_pred_gather(const uint32_t *tile_count, const int64_t *tile_prefix, const int64_t *slots,
                              int64_t ntiles, int64_t *pos)
{ const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const uint32_t c = min(tile_count[t],(uint32_t) kSlot);
  if (c == 0) return;
  int64_t p[kSlot];
  for (uint32_t i = 0; i < c; i++)                       // insertion sort of at most kSlot positions
    { const int64_t x = slots[(size_t) t * kSlot + i];
      int k = (int) i;
      while (k > 0 && p[k-1] > x) { p[k] = p[k-1]; k--; }
      p[k] = x;
    }
  int64_t *dst = pos + tile_prefix[t];
  for (uint32_t i = 0; i < c; i++) dst[i] = p[i];
}

template <int PRED>
int index_positions(dx_ctx *ctx, const uint8_t *buf, size_t n, size_t first,
                    int64_t **d_pos, int64_t *count)
{ const int64_t ntiles = (int64_t) ((n + kTileBytes - 1) / kTileBytes);
  *d_pos = NULL; *count = 0;
  if (ntiles == 0) return DX_OK;
  if (ctx->route[DXR_EXACT_INDEX]) return index_positions_exact<PRED>(ctx,buf,n,first,d_pos,count);
  uint32_t *d_cnt  = (uint32_t *) dx_arena_get(ctx,(size_t) ntiles*4 + 16);
  int64_t  *d_pre  = (int64_t *)  dx_arena_get(ctx,(size_t) (ntiles+1)*8);
  int64_t  *d_slot = (int64_t *)  dx_arena_get(ctx,(size_t) ntiles*kSlot*8);
  if (d_cnt == NULL || d_pre == NULL || d_slot == NULL) return DX_E_NOMEM;
  int32_t *d_over = (int32_t *) (d_cnt + ntiles);
  DX_CUDA(ctx,cudaMemsetAsync(d_over,0,4,ctx->stream));
  if (PRED == DX_PRED_NEWLINE && ctx->route[DXR_INDEX_BULK])
    { const size_t smem = 2*(size_t) kTileBytes;
      DX_CUDA(ctx,cudaFuncSetAttribute(k_pred_slots_bulk<PRED>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem));
      int64_t grid = (int64_t) ctx->sm_count * (ctx->route[DXR_INDEX_BULK] > 1 ? ctx->route[DXR_INDEX_BULK] : 4);
      if (grid > ntiles) grid = ntiles;
      DX_PROF_BEGIN(ctx);
      k_pred_slots_bulk<PRED><<<(unsigned) grid,kTileThreads,smem,ctx->stream>>>(buf,n,first,ntiles,d_cnt,d_slot,d_over);
      DX_LAUNCHED(ctx,"k_pred_slots_bulk");
    }
  else
    { DX_PROF_BEGIN(ctx);
      k_pred_slots<PRED><<<(unsigned) ntiles,kTileThreads,0,ctx->stream>>>(buf,n,first,d_cnt,d_slot,d_over);
      DX_LAUNCHED(ctx,"k_pred_slots");
    }
  { const int rc = launch_scan(ctx,d_cnt,ntiles,d_pre,"k_tile_scan");
    if (rc != DX_OK) return rc;
  }
  struct Res { int64_t total; int32_t over; int32_t pad; };
  Res *hp = (Res *) dx_hpin_get(ctx,sizeof(Res));
  if (hp == NULL) return DX_E_NOMEM;
  { int rc;
    if ((rc = dxk_fetch(ctx,&hp->total,d_pre+ntiles,8)) != DX_OK) return rc;
    if ((rc = dxk_fetch(ctx,&hp->over,d_over,4)) != DX_OK) return rc;
  }
  if (ctx->overlap_fn != NULL)                  // the caller's host work, hidden behind the index pass
    { void (*fn)(void *) = ctx->overlap_fn;
      ctx->overlap_fn = NULL;
      fn(ctx->overlap_arg);
    }
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  const Res h = *hp;
  if (h.over) return index_positions_exact<PRED>(ctx,buf,n,first,d_pos,count);
  *count = h.total;
  if (h.total == 0) return DX_OK;
  int64_t *pos = (int64_t *) dx_arena_get(ctx,(size_t) h.total*8);
  if (pos == NULL) return DX_E_NOMEM;
  DX_PROF_BEGIN(ctx);
  k_pred_gather<<<(unsigned) ((ntiles + 255)/256),256,0,ctx->stream>>>(d_cnt,d_pre,d_slot,ntiles,pos);
  DX_LAUNCHED(ctx,"k_pred_gather");
  *d_pos = pos;
  return DX_OK;
}

// For every candidate field position q: how many 0xff bytes sit directly before byte q-1 (capped
// at 2^20), what byte q-1 is, and the field bytes at q.  Lets the host verify "previous entry
// ended at p, this entry's well-delta bytes are exactly [p, q)" and rebuild the header text
// without seeing the image (dexqv.c:128-139).
__global__ void k_cand_context(const uint8_t *buf, size_t n, size_t first, const int64_t *q,
                               int64_t count, int fieldbytes, CandInfo *info)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  CandInfo ci;
  const int64_t p = q[i] - 1;                      // the delta terminator byte
  for (int k = 0; k < 16; k++)
    ci.field[k] = (k < fieldbytes && (size_t) (q[i] + k) < n) ? buf[q[i] + k] : 0;
  ci.pad[0] = ci.pad[1] = ci.pad[2] = 0;
  if (p < (int64_t) first)
    { ci.ffrun = -1; ci.last = 0; }
  else
    { ci.last = buf[p];
      int32_t r = 0;
      int64_t k = p - 1;
      while (k >= (int64_t) first && buf[k] == 0xff && r < (1 << 20)) { r++; k--; }
      ci.ffrun = r;
    }
  info[i] = ci;
}

__global__ void k_skip_ff(const uint8_t *buf, int64_t n, const int64_t *start, int64_t count, int64_t *q)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int64_t p = start[i];
  while (p < n && buf[p] == 0xff) p++;
  q[i] = p + 1;
}

__global__ void k_field_rlen(const uint8_t *buf, const int64_t *q, int64_t count, int32_t *rlen)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint8_t *p = buf + q[i];
  rlen[i] = (int32_t) (load_le32(p+4) - load_le32(p));
}


// ---- ticket order ---------------------------------------------------------------------------------
// The per-entry kernels hand out work through a ticket counter; whatever is handed out last is the
// tail of the launch, so tickets go to the longest entries first: a counting sort on rlen / 512
// (any order inside a bucket).  One CTA; N is tens of thousands.
constexpr int kOrderBuckets = 2048;              // rlen / 256, clamped: popular lengths spread over many counters

__device__ __forceinline__ int order_bucket(int32_t rl)
{ const int q = (rl <= 0) ? 0 : min(kOrderBuckets - 1,rl >> 8);
  return kOrderBuckets - 1 - q;                    // descending length
}

__global__ void __launch_bounds__(1024)
k_ticket_order(const int32_t *rlen, int64_t n, int32_t *order)
{ __shared__ uint32_t cnt[kOrderBuckets];
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = threadIdx.x; b < kOrderBuckets; b += 1024) cnt[b] = 0;
  __syncthreads();
  // (eight independent loads in flight per thread: the kernel is one CTA and bound by their latency)
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += 8*1024)
    { int32_t r[8];
#pragma unroll
      for (int k = 0; k < 8; k++) r[k] = (i0 + k*1024 < n) ? rlen[i0 + k*1024] : -1;
#pragma unroll
      for (int k = 0; k < 8; k++) if (i0 + k*1024 < n) atomicAdd(&cnt[order_bucket(r[k])],1u);
    }
  __syncthreads();
  // exclusive scan of the counters: two per thread
  const uint32_t c0 = cnt[2*threadIdx.x], c1 = cnt[2*threadIdx.x + 1];
  const uint32_t inc = dx_warp_incl_sum(c0 + c1,lane);
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0)
    { const uint32_t w = wsum[lane];
      const uint32_t wi = dx_warp_incl_sum(w,lane);
      wsum[lane] = wi - w;
    }
  __syncthreads();
  const uint32_t excl = wsum[warp] + inc - (c0 + c1);
  cnt[2*threadIdx.x] = excl; cnt[2*threadIdx.x + 1] = excl + c0;
  __syncthreads();
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += 8*1024)
    { int32_t r[8];
#pragma unroll
      for (int k = 0; k < 8; k++) r[k] = (i0 + k*1024 < n) ? rlen[i0 + k*1024] : -1;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (i0 + k*1024 < n) order[atomicAdd(&cnt[order_bucket(r[k])],1u)] = (int32_t) (i0 + k*1024);
    }
}

}  // namespace

int dxk_ticket_order(dx_ctx *ctx, const int32_t *d_rlen, int64_t n, int32_t *d_order)
{ if (n == 0) return DX_OK;
  DX_PROF_BEGIN(ctx); k_ticket_order<<<1,1024,0,ctx->stream>>>(d_rlen,n,d_order);
  DX_LAUNCHED(ctx,"k_ticket_order");
  return DX_OK;
}

int dxk_scan_u32(dx_ctx *ctx, const uint32_t *d_in, int64_t n, int64_t *d_prefix)
{ return launch_scan(ctx,d_in,n,d_prefix,"k_scan_u32");
}

int dxk_cand_context(dx_ctx *ctx, const uint8_t *d_in, size_t n, size_t first, const int64_t *d_q,
                     int64_t count, int fieldbytes, CandInfo *d_info)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx); k_cand_context<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_in,n,first,d_q,count,fieldbytes,d_info);
  DX_LAUNCHED(ctx,"k_cand_context");
  return DX_OK;
}

int dxk_skip_ff(dx_ctx *ctx, const uint8_t *d_in, size_t n, const int64_t *d_start, int64_t count,
                int64_t *d_q)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx); k_skip_ff<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_in,(int64_t) n,d_start,count,d_q);
  DX_LAUNCHED(ctx,"k_skip_ff");
  return DX_OK;
}

int dxk_field_rlen(dx_ctx *ctx, const uint8_t *d_in, const int64_t *d_q, int64_t count, int32_t *d_rlen)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx); k_field_rlen<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_in,d_q,count,d_rlen);
  DX_LAUNCHED(ctx,"k_field_rlen");
  return DX_OK;
}

// NOTE: `buf` must be readable up to the next 16-byte multiple of n (all library-owned and
// torch-allocated device buffers are; dx_api pads the ones it stages itself).
int dxk_index_positions(dx_ctx *ctx, int pred, const uint8_t *d_buf, size_t n, size_t first,
                        int64_t **d_pos, int64_t *count)
{ switch (pred)
    { case DX_PRED_NEWLINE:   return index_positions<DX_PRED_NEWLINE>(ctx,d_buf,n,first,d_pos,count);
      case DX_PRED_FASTA_HDR: return index_positions<DX_PRED_FASTA_HDR>(ctx,d_buf,n,first,d_pos,count);
      case DX_PRED_QVCAND:    return index_positions<DX_PRED_QVCAND>(ctx,d_buf,n,first,d_pos,count);
      case DX_PRED_ARCAND:    return index_positions<DX_PRED_ARCAND>(ctx,d_buf,n,first,d_pos,count);
    }
  return dx_fail(ctx,DX_E_ARG,"unknown position predicate %d",pred);
}

int dxk_qv_entries(dx_ctx *ctx, const uint8_t *d_text, size_t n, const int64_t *d_nl,
                   int64_t nlines, QvEntries ent, int32_t *h_err, uint64_t *h_totchar,
                   int64_t *h_noncanon, int64_t *h_last_nl)
{ (void) n;
  const int64_t nent = nlines / 6;
  h_err[0] = 0; h_err[1] = 0; *h_totchar = 0; *h_noncanon = 0; *h_last_nl = -1;
  if (nent == 0) return DX_OK;
  unsigned long long *d_err = (unsigned long long *) dx_arena_get(ctx,32);
  unsigned long long *res = (unsigned long long *) dx_hpin_get(ctx,32);
  if (d_err == NULL || res == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_err,0xff,8,ctx->stream));
  DX_CUDA(ctx,cudaMemsetAsync(d_err+1,0,16,ctx->stream));
  DX_CUDA(ctx,cudaMemsetAsync(d_err+3,0xff,8,ctx->stream));
  DX_PROF_BEGIN(ctx); k_qv_entries<<<(unsigned) ((nent+255)/256),256,0,ctx->stream>>>(d_text,d_nl,nent,ent,d_err);
  DX_LAUNCHED(ctx,"k_qv_entries");
  DX_CUDA(ctx,cudaMemcpyAsync(res,d_err,32,cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  const unsigned long long e = res[0];
  *h_totchar = res[1];
  *h_noncanon = (int64_t) res[2];
  *h_last_nl = (int64_t) res[3];
  if (e != ~0ull)
    { h_err[0] = (int32_t) (e & 0xff);
      h_err[1] = (int32_t) (e >> 8);        // 1-based line number
    }
  return DX_OK;
}
