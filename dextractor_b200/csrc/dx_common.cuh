// dx_common.cuh -- device helpers shared by the sm_100a kernels of libdexb200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define DX_FULL 0xffffffffu

// ---- streaming global loads ------------------------------------------------------------------
// The text / compressed images are read once per pass: read-only path, do not pollute L1.
__device__ __forceinline__ uint4 dx_ldg16(const void *p)            // p must be 16-byte aligned
{ uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void dx_stg16(void *p, uint4 v)          // p must be 16-byte aligned
{ asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 16 bytes at any address (two aligned loads + funnel shifts); never reads past text_end16
__device__ __forceinline__ uint4 dx_ld16_any(const uint8_t *p, const uint8_t *end16)
{ const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint8_t *al = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t) 15);
  const int sk = (int) (a & 15);
  uint4 lo = dx_ldg16(al);
  if (sk == 0) return lo;
  uint4 hi = (al + 16 < end16) ? dx_ldg16(al + 16) : make_uint4(0,0,0,0);
  const uint32_t sb = (sk & 3) * 8;
  uint4 r;
  switch (sk >> 2)
    { case 0:  r.x = __funnelshift_r(lo.x,lo.y,sb); r.y = __funnelshift_r(lo.y,lo.z,sb);
               r.z = __funnelshift_r(lo.z,lo.w,sb); r.w = __funnelshift_r(lo.w,hi.x,sb); break;
      case 1:  r.x = __funnelshift_r(lo.y,lo.z,sb); r.y = __funnelshift_r(lo.z,lo.w,sb);
               r.z = __funnelshift_r(lo.w,hi.x,sb); r.w = __funnelshift_r(hi.x,hi.y,sb); break;
      case 2:  r.x = __funnelshift_r(lo.z,lo.w,sb); r.y = __funnelshift_r(lo.w,hi.x,sb);
               r.z = __funnelshift_r(hi.x,hi.y,sb); r.w = __funnelshift_r(hi.y,hi.z,sb); break;
      default: r.x = __funnelshift_r(lo.w,hi.x,sb); r.y = __funnelshift_r(hi.x,hi.y,sb);
               r.z = __funnelshift_r(hi.y,hi.z,sb); r.w = __funnelshift_r(hi.z,hi.w,sb); break;
    }
  return r;
}

// ---- SWAR byte predicates on a 32-bit word (4 text bytes) --------------------------------------
// 0x80 in every byte lane of x that equals c (exact, no cross-lane borrow).
__device__ __forceinline__ uint32_t dx_eq_mask(uint32_t x, uint32_t c)
{ uint32_t t = x ^ (c * 0x01010101u);
  return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
}

// gather the four 0x80 flags of a SWAR mask into bits 0..3 (byte 0 -> bit 0)
__device__ __forceinline__ uint32_t dx_nibble(uint32_t m)
{ return (((m >> 7) * 0x00204081u) >> 21) & 0xfu;
}

// 16-bit mask (bit i = byte i of the 16-byte chunk equals c)
__device__ __forceinline__ uint32_t dx_eq_mask16(uint4 v, uint32_t c)
{ return  dx_nibble(dx_eq_mask(v.x,c))        | (dx_nibble(dx_eq_mask(v.y,c)) << 4)
       | (dx_nibble(dx_eq_mask(v.z,c)) << 8)  | (dx_nibble(dx_eq_mask(v.w,c)) << 12);
}

// byte i (0..15) of a 16-byte chunk held in registers, i not a compile-time constant
__device__ __forceinline__ uint32_t dx_byte_of(uint4 v, int i)
{ uint32_t w = (i & 8) ? ((i & 4) ? v.w : v.z) : ((i & 4) ? v.y : v.x);
  return (w >> ((i & 3) * 8)) & 0xffu;
}

// bits [lo,hi) set, 0 <= lo <= hi <= 16
__device__ __forceinline__ uint32_t dx_range16(int lo, int hi)
{ return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}

// ---- warp collectives --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dx_warp_incl_sum(uint32_t v, int lane)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
    { uint32_t t = __shfl_up_sync(DX_FULL, v, d);
      if (lane >= d) v += t;
    }
  return v;
}

__device__ __forceinline__ uint64_t dx_warp_incl_sum64(uint64_t v, int lane)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
    { uint64_t t = __shfl_up_sync(DX_FULL, v, d);
      if (lane >= d) v += t;
    }
  return v;
}

__device__ __forceinline__ int dx_warp_incl_max(int v, int lane)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
    { int t = __shfl_up_sync(DX_FULL, v, d);
      if (lane >= d) v = max(v, t);
    }
  return v;
}

__device__ __forceinline__ uint32_t dx_warp_sum(uint32_t v)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    v += __shfl_xor_sync(DX_FULL, v, d);
  return v;
}

// ---- staged bytes (shared memory, 4-byte aligned, one readable pad word after the data)
//      -> global memory at ANY byte alignment, by one warp --------------------------------------
// Interior destination words are written as aligned 32-bit stores assembled with a funnel
// shift; the ragged head and tail use byte stores so neighbouring units (other warps / CTAs)
// that own the other bytes of the same word are never clobbered.
__device__ __forceinline__ void dx_warp_copy_out(uint8_t *gdst, const uint32_t *ssrc, uint32_t n,
                                                 int lane)
{ const uint8_t *sb = reinterpret_cast<const uint8_t *>(ssrc);
  uint32_t head = (4u - (uint32_t) (reinterpret_cast<uintptr_t>(gdst) & 3u)) & 3u;
  if (head > n) head = n;
  if ((uint32_t) lane < head)
    gdst[lane] = sb[lane];
  const uint32_t body = (n - head) >> 2;
  uint32_t *gw = reinterpret_cast<uint32_t *>(gdst + head);
  const uint32_t sh = head * 8u;                   // source is `head` bytes ahead of a word
  for (uint32_t i = lane; i < body; i += 32)
    { uint32_t lo = ssrc[i], hi = ssrc[i+1];
      gw[i] = __funnelshift_r(lo, hi, sh);        // sh == 0 -> lo
    }
  const uint32_t done = head + 4u*body;
  if ((uint32_t) lane < n - done)
    gdst[done + lane] = sb[done + lane];
}

// ---- bit append into a zeroed shared-memory word array, MSB first ------------------------------
// `pos` is the bit offset from word 0 of `stage`; len in 1..32 (code must fit in len bits).
__device__ __forceinline__ void dx_or_bits(uint32_t *stage, uint32_t pos, uint32_t code, uint32_t len)
{ const uint32_t w = pos >> 5, off = pos & 31u;
  if (off + len <= 32u)
    atomicOr(&stage[w], code << (32u - off - len));
  else
    { const uint32_t spill = off + len - 32u;
      atomicOr(&stage[w], code >> spill);
      atomicOr(&stage[w+1], code << (32u - spill));
    }
}
