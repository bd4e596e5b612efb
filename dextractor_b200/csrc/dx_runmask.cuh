// dx_runmask.cuh -- run-length histogram of a del / sub line from bit masks.
//
// NOT on the product path yet: this is the arithmetic of the planned replacement of k_qv_hist_run's
// per-item queue, written so that the same code compiles for the host (tests/hostfuzz/fz_runmask.cpp
// checks it against a direct restatement of Histogram_Runs, QV.c:709-724) and for the device.
//
// The line is cut into 32-byte spans.  For a span, C has bit i set when byte i is an ITEM (inside
// the line and not the run character), B are its boundaries (= C, plus the line start) and P the
// boundaries of the previous span.  An item's run length g is the distance to the previous boundary
// minus one.  With D_k = the boundary mask delayed by k positions (one funnel shift of P:B), the
// items with run length g are  Z_g & D_{g+1}  where Z_0 = C and Z_{g+1} = Z_g & ~D_{g+1}: five
// operations per g for 32 bytes, no per-item work.  Whatever is left in Z_32 are items whose run is
// longer than 31; at most one per span (its first item), resolved from the position of the last
// boundary before the span (on the device: a warp max-scan).
//
// Line edges (QV.c:713-722):
//   * the run before the first item starts at the line start: position -1 is a boundary (not an item),
//     the last bit of P when the line starts on the span lattice, else a bit of the first span's B;
//   * a trailing run (the line does not end in an item) is counted like an item at position rlen;
//     if rlen-1 is an item there is no trailing run (it would be a run of length 0, which the
//     reference does not count there).
#ifndef DX_RUNMASK_CUH
#define DX_RUNMASK_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define DX_HD __host__ __device__ __forceinline__
#else
#define DX_HD static inline
#endif

// bit i of the result = bit (i - k) of the 64-bit string prev:cur (cur in the high half), 0 <= k <= 32
DX_HD uint32_t dx_delay_mask(uint32_t prev, uint32_t cur, int k)
{
#if defined(__CUDA_ARCH__)
  return (k == 32) ? prev : __funnelshift_l(prev,cur,k);
#else
  const uint64_t w = ((uint64_t) cur << 32) | prev;
  return (uint32_t) (w >> (32 - k));
#endif
}

DX_HD int dx_popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// cnt[g] += number of the items C of the span whose run length is g, g = 0..31.  B are the span's
// boundaries (the items, plus position -1 when it falls into the span), P the previous span's.
// -> mask of the items whose run length is 32 or more (0 or the span's first item).
DX_HD uint32_t dx_gap_counts32(uint32_t C, uint32_t B, uint32_t P, uint32_t cnt[32])
{ uint32_t Z = C;
#pragma unroll
  for (int g = 0; g < 32; g++)
    { const uint32_t D = dx_delay_mask(P,B,g+1);
      cnt[g] += (uint32_t) dx_popc32(Z & D);
      Z &= ~D;
    }
  return Z;
}

#if !defined(__CUDA_ARCH__)
// Host model of the whole-line procedure the kernel will follow (one lane = one span): run[] gets
// exactly what Histogram_Runs(run,line,rlen,rc) adds.  `skew` shifts the span lattice against the
// line start, as the 32-byte alignment of a line in the text image does (0 <= skew < 32).
static inline void dx_runs_line_model(uint64_t run[256], const uint8_t *line, int rlen, int rc, int skew)
{ if (rlen <= 0) return;
  const int trailing = (line[rlen-1] == (uint8_t) rc);          // a virtual item at rlen closes the line
  const long nitem = (long) rlen + trailing;                     // positions that can hold an item
  uint32_t P = (skew == 0) ? 0x80000000u : 0u;                   // position -1 = last bit of the span before
  long lastb = -1;                                               // last boundary before the span
  for (long s0 = -skew; s0 < nitem; s0 += 32)                    // the span covers positions s0 .. s0+31
    { uint32_t C = 0;
      for (int i = 0; i < 32; i++)
        { const long p = s0 + i;
          if (p < 0 || p >= nitem) continue;
          if (p == rlen || line[p] != (uint8_t) rc) C |= 1u << i;
        }
      uint32_t B = C;
      if (s0 < 0) B |= 1u << (skew - 1);                         // position -1 lies in the first span
      uint32_t cnt[32] = { 0 };
      const uint32_t Z = dx_gap_counts32(C,B,P,cnt);
      for (int g = 0; g < 32; g++) run[g] += cnt[g];
      if (Z)                                                     // the span's first item, run >= 32
        { const long gap = (s0 + __builtin_ctz(Z)) - lastb - 1;
          run[gap >= 255 ? 255 : gap] += 1;
        }
      if (B) lastb = s0 + 31 - __builtin_clz(B);
      P = B;
    }
}
#endif

#endif
