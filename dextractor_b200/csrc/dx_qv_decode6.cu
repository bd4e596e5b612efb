// dx_qv_decode6.cu -- .dexqv entry decoder, one LANE per entry (the library's main decoder).
//
// Replaces Decode / Decode_Run (reference QV.c:510-691) + Packed_Length / Unpack_Tag
// (QV.c:823-847) + the per-entry text output of undexqv.c:182-207.
//
// A Huffman stream stores no code boundaries, so a stream can only be split among threads by
// speculation (dx_qv_decode5.cu: every symbol is decoded 2-3 times and a third of the lanes idle).
// But a 2 GB file holds ~40 000 independent entries: here every lane of a warp decodes ONE WHOLE
// ENTRY sequentially -- exactly the reference's loop, one table lookup per one or two symbols, no
// speculation, no warp collective anywhere in the kernel.  The 32 entries of a warp are neighbours
// in the longest-first ticket order, so their lengths agree within a few percent and the lanes
// stay busy together.  Entries too long for this (a lane needs ~100 cycles per position) go to the
// warp-per-entry kernel; the host picks the cut (dxk_qv_decode6x).
//
// What makes a sequential lane fast enough:
//   * decode tables for all four streams resident in shared memory (64 KB per CTA, 12-bit index,
//     two symbols per lookup on the plain streams);
//   * the lane's bit window is a left-aligned 64-bit register pair; it is refilled one 32-bit word
//     at a time from a 16-word per-lane ring in shared memory ([word][lane]: conflict free), and
//     the ring is topped up with 16-byte global loads issued two quads ahead (software pipelined:
//     the lane never waits for HBM).  The byte misalignment of a stream is removed when a quad
//     enters the ring, so the hot loop sees aligned stream words;
//   * output goes through a 64-byte per-lane ring in shared memory (stride 68 bytes: lanes at the
//     same offset hit different banks) and leaves in aligned 16-byte stores; the five lines of an
//     entry are contiguous in the text, so one byte stream per lane runs through all of them and only
//     the first and last block of an entry are written bytewise.  Run-length streams keep the ring
//     pre-filled with the run character: a run is an addition to the write pointer;
//   * rare per-lane events (ring top-up, block flush) are polled every 8 lookups, not every lookup:
//     what is rare for a lane happens in almost every iteration of a 32-lane warp.

#include <stdio.h>
#include <stdlib.h>
#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kWarps6   = 8;
constexpr int kThreads6 = kWarps6 * 32;
constexpr int kRingW    = 16;                   // stream words per lane in the input ring
constexpr int kOutRing  = 64;                   // bytes per lane in the output ring
constexpr int kOutPitch = 68;                   // ... and its pitch (17 words: odd, so conflict free)

struct Dec6Args
{ const uint8_t *in;
  int64_t        n;
  const QvDecTables4 *tab;
  int32_t        delchar, subchar, upper, write;   // write: 1 text + header, 2 lines only (speculative,
                                                   //        per-entry status)
  int64_t        first, count;                     // tickets [first, count)
  const int64_t *start;        // first stream byte of each entry (after beg/end/qv)
  const int32_t *rlen;
  const QvDecEntry *ent;       // output placement (write == 1, or write == 2 with ent)
  const int64_t *toff;         // write == 2 without ent: where the entry's lines go
  const char    *prefix; int32_t plen;
  uint8_t       *out;
  int64_t       *soff;         // [count][6] or NULL
  int32_t       *status;       // [count] (write == 2) or [1] (write == 1)
  const int64_t *limit;        // per entry: first byte the entry may not reach (NULL: the image end)
  const int32_t *order;        // ticket -> entry (long entries first), NULL: identity
};

struct Shared6
{ uint32_t tab[4][4096];                        // del, ins, mrg, sub: one multi table or run|sym u16 tables
  uint32_t ring[kWarps6][kRingW][32];
  uint32_t outr[kWarps6][32*kOutPitch/4];
};

extern __shared__ __align__(16) uint8_t dx_dec6_smem[];

// ---- the lane's input: a bit window over a byte-aligned stream of little-endian 32-bit words ----
struct BitIn
{ uint32_t hi, lo;             // the next `valid` stream bits, left aligned in hi:lo
  int32_t  valid;
  uint32_t cons;               // bits consumed since the stream's start
  uint32_t rd, st;             // ring: next word to read / words stored (indices of aligned words)
  uint32_t sb8;                // byte skew of the stream against the aligned words, in bits
  const uint4 *src, *end16;    // next quad to load / first quad that may not be read
  uint4    pa, pb;             // quads st/4 and st/4 + 1 (pb may still be in flight)
  uint32_t *ring;              // &ring[warp][0][lane]
  uint64_t base;               // address of aligned word 0

  __device__ __forceinline__ uint4 load(const uint4 *p) const
  { return (p < end16) ? __ldg(p) : make_uint4(0,0,0,0); }

  // aligned quad pa (its successor's first word is pb.x) -> four stream-aligned words in the ring
  __device__ __forceinline__ void store_quad()
  { uint32_t *r = ring + (st & (kRingW-1))*32;
    r[0]  = __funnelshift_r(pa.x,pa.y,sb8);
    r[32] = __funnelshift_r(pa.y,pa.z,sb8);
    r[64] = __funnelshift_r(pa.z,pa.w,sb8);
    r[96] = __funnelshift_r(pa.w,pb.x,sb8);
    st += 4;
    pa = pb;
    pb = load(src); src++;
  }

  // at least 9 words buffered ahead of rd (a poll interval consumes at most 7)
  __device__ __forceinline__ void top_up()
  { while (st - rd <= 8u) store_quad(); }

  template <bool SWAP>         // SWAP: the stream is a byte string, first byte first (packed tags)
  __device__ __forceinline__ void refill()
  { if (valid <= 32)
      { uint32_t x = ring[(rd & (kRingW-1))*32];
        if (SWAP) x = __byte_perm(x,0,0x0123);
        rd++;
        hi |= __funnelshift_rc(x,0u,(uint32_t) valid);
        lo  = __funnelshift_lc(0u,x,32u - (uint32_t) valid);
        valid += 32;
      }
  }

  template <bool SWAP>
  __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *image_end, uint32_t *ring_lane)
  { const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    ring = ring_lane;
    src = reinterpret_cast<const uint4 *>(a & ~(uintptr_t) 15);
    end16 = reinterpret_cast<const uint4 *>((reinterpret_cast<uintptr_t>(image_end) + 15) & ~(uintptr_t) 15);
    base = (uint64_t) (a & ~(uintptr_t) 15);
    sb8 = (uint32_t) (a & 3) * 8u;
    rd = (uint32_t) (a & 15) >> 2; st = 0;
    const uint4 q0 = load(src), q1 = load(src+1), q2 = load(src+2), q3 = load(src+3);
    pa = q0; pb = q1; src += 2;
    // (store_quad loads the quad after pb itself: unroll the first three by hand so that the four
    //  loads above are in flight together)
    { uint32_t *r = ring;
      r[0]   = __funnelshift_r(q0.x,q0.y,sb8); r[32]  = __funnelshift_r(q0.y,q0.z,sb8);
      r[64]  = __funnelshift_r(q0.z,q0.w,sb8); r[96]  = __funnelshift_r(q0.w,q1.x,sb8);
      r[128] = __funnelshift_r(q1.x,q1.y,sb8); r[160] = __funnelshift_r(q1.y,q1.z,sb8);
      r[192] = __funnelshift_r(q1.z,q1.w,sb8); r[224] = __funnelshift_r(q1.w,q2.x,sb8);
      r[256] = __funnelshift_r(q2.x,q2.y,sb8); r[288] = __funnelshift_r(q2.y,q2.z,sb8);
      r[320] = __funnelshift_r(q2.z,q2.w,sb8); r[352] = __funnelshift_r(q2.w,q3.x,sb8);
    }
    st = 12; pa = q3; src = reinterpret_cast<const uint4 *>(a & ~(uintptr_t) 15) + 4;
    pb = load(src); src++;
    hi = 0; lo = 0; valid = 0; cons = 0;
    refill<SWAP>(); refill<SWAP>();
  }

  __device__ __forceinline__ void take(uint32_t len)           // len < 32, len <= valid
  { hi = __funnelshift_l(lo,hi,len);
    lo <<= len;
    valid -= (int32_t) len;
    cons += len;
  }

  // first byte of the image that has not been handed to the window yet
  __device__ __forceinline__ uint64_t reached() const { return base + (uint64_t) rd*4u; }
};

// ---- the lane's output: bytes appended to one contiguous span of the text -------------------------
struct OutSt
{ uint8_t *gbase;              // 16-byte aligned; byte counter t <-> gbase + t
  uint32_t wr, fl, head;       // appended / flushed (multiple of 16) / bytes of block 0 that are not ours
  uint8_t *ring;               // this lane's 64 bytes of shared memory

  __device__ __forceinline__ void init(uint8_t *dst, uint8_t *ring_lane)
  { const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    gbase = reinterpret_cast<uint8_t *>(a & ~(uintptr_t) 15);
    head = (uint32_t) (a & 15); wr = head; fl = 0; ring = ring_lane;
  }
  __device__ __forceinline__ void put(uint32_t c) { ring[wr & (kOutRing-1)] = (uint8_t) c; wr++; }

  // FILL != 0: a flushed block is filled with the byte again (run-length streams)
  template <bool FILL>
  __device__ __forceinline__ void flush(uint32_t fill4)
  { while (wr - fl >= 16u)
      { uint32_t *b = reinterpret_cast<uint32_t *>(ring + (fl & (kOutRing-1)));
        const uint4 v = make_uint4(b[0],b[1],b[2],b[3]);
        if (fl == 0 && head != 0)
          { const uint8_t *bb = reinterpret_cast<const uint8_t *>(b);
            for (uint32_t k = head; k < 16u; k++) gbase[k] = bb[k];
          }
        else
          dx_stg16(gbase + fl,v);
        if (FILL) { b[0] = fill4; b[1] = fill4; b[2] = fill4; b[3] = fill4; }
        fl += 16;
      }
  }
  // every byte appended so far is in global memory afterwards; the state does not change
  __device__ __forceinline__ void sync_partial()
  { flush<false>(0);
    const uint32_t from = (fl == 0) ? head : fl;
    for (uint32_t t = from; t < wr; t++) gbase[t] = ring[t & (kOutRing-1)];
  }
  // bytes [wr, fl + 64) of the ring := c
  __device__ __forceinline__ void prefill(uint32_t c)
  { for (uint32_t t = wr; t < fl + kOutRing; t++) ring[t & (kOutRing-1)] = (uint8_t) c; }

  // append r copies of the byte the ring is pre-filled with
  __device__ __forceinline__ void skip(uint32_t r, uint32_t fill4)
  { while (true)
      { const uint32_t room = fl + 48u - wr;           // >= 33: flush leaves fewer than 16 pending
        const uint32_t a = min(r,room);
        wr += a; r -= a;
        flush<true>(fill4);
        if (r == 0) break;
      }
  }
};

// codes longer than 12 bits (same tables as the warp-per-entry kernel)
__device__ __forceinline__ uint32_t lookup_long6(const QvDecTables2 *t, int k, uint32_t w16)
{ uint32_t e = __ldg(&t->prim[k][w16 >> 5]);
  if (e & 0x8000u)
    e = __ldg(&t->sub[k][(e & 0x7fffu)*32u + (w16 & 31u)]);
  return e;
}
__device__ __noinline__ uint32_t long_entry6(const QvDecTables2 *t, int k, uint32_t w, uint32_t *bad)
{ const uint32_t f = lookup_long6(t,k,w >> 16);
  uint32_t len = (f >> 8) & 31u;
  const uint32_t c = f & 0xffu;
  if (len == 0) { len = 1; *bad = 1; }
  if (t->type[k] == 2 && c == 255u)
    return (len + 8u) | (1u << 5) | 0x80u | (len << 8) | (255u << 16);
  return len | (1u << 5) | (len << 8) | (c << 16);
}
__device__ __noinline__ uint32_t long_single6(const QvDecTables2 *t, int k, uint32_t w, uint32_t *bad)
{ uint32_t f = lookup_long6(t,k,w >> 16) & 0x1fffu;
  if ((f >> 8) == 0u) { f |= 0x100u; *bad = 1; }
  return f;
}

// multi entry (QvDecTables4): bits 0-4 total length, 5-6 symbols, 7 escape, 8-12 length of the first
// code, 16-23 / 24-31 the symbols
#define E6_LEN(e)   ((e) & 31u)
#define E6_N(e)     (((e) >> 5) & 3u)
#define E6_LEN0(e)  (((e) >> 8) & 31u)

struct Lane
{ BitIn in; OutSt out;
  uint32_t bad;
  uint64_t lim;                // address the entry may not read past
};

// one plain stream of L symbols (QV.c:510-599); returns the bytes the stream occupies in the file
__device__ __forceinline__ uint32_t plain_stream(Lane &ln, const uint32_t *mt, const QvDecTables2 *t2, int symtab,
                                                 uint32_t L, bool wr, uint32_t *kept_all)
{ BitIn &in = ln.in; OutSt &out = ln.out;
  uint32_t cnt = 0, last = 0;
  // two symbols per lookup while at least three are still to come
  while (cnt + 2u < L)
    {
#pragma unroll 1
      for (int it = 0; it < 8 && cnt + 2u < L; it++)
        { const uint32_t w = in.hi;
          uint32_t e = mt[w >> 20];
          if (e == 0u) e = long_entry6(t2,symtab,w,&ln.bad);
          uint32_t c0 = (e >> 16) & 0xffu;
          if (e & 0x80u) c0 = (w << E6_LEN0(e)) >> 24;            // the literal after the escape
          if (wr)
            { out.ring[out.wr & (kOutRing-1)] = (uint8_t) c0;
              out.ring[(out.wr + 1u) & (kOutRing-1)] = (uint8_t) (e >> 24);
              out.wr += E6_N(e);
            }
          cnt += E6_N(e);
          in.take(E6_LEN(e));
          in.refill<false>();
        }
      in.top_up();
      if (wr) out.flush<false>(0);
      if (in.reached() > ln.lim) { ln.bad = 1; return 0; }
    }
  // the last one or two symbols one at a time: `last` is the position of the last item (the
  // literal of an escape), which fixes the stream's length in the file (QV.c:537-551)
  while (cnt < L)
    { const uint32_t w = in.hi;
      uint32_t e = mt[w >> 20];
      if (e == 0u) e = long_entry6(t2,symtab,w,&ln.bad);
      uint32_t c0 = (e >> 16) & 0xffu;
      const uint32_t l0 = E6_LEN0(e);
      uint32_t adv = l0;
      last = in.cons;
      if (e & 0x80u) { c0 = (w << l0) >> 24; last += l0; adv += 8u; }
      if (wr) out.put(c0);
      cnt++;
      in.take(adv);
      in.refill<false>();
    }
  if (kept_all != NULL) *kept_all = L;
  return (L == 0) ? 0u : ((last + 47u) >> 5) * 4u;
}

// one run-length stream (QV.c:604-691): (run of rc, one other symbol)*
__device__ __forceinline__ uint32_t run_stream(Lane &ln, const uint16_t *rt, const uint16_t *st,
                                               const QvDecTables2 *t2, int symtab, int runtab, bool esc,
                                               uint32_t rc, uint32_t L, bool wr, uint32_t *kept_out)
{ BitIn &in = ln.in; OutSt &out = ln.out;
  const uint32_t rc4 = rc * 0x01010101u;
  uint32_t cnt = 0, last = 0, kept = 0;
  if (wr) out.prefill(rc);
  while (cnt < L)
    {
#pragma unroll 1
      for (int it = 0; it < 4 && cnt < L; it++)
        { uint32_t w = in.hi;
          uint32_t e = rt[w >> 20];
          if (e == 0u) e = long_single6(t2,runtab,w,&ln.bad);
          uint32_t r = e & 0xffu;
          last = in.cons;
          in.take(e >> 8);
          in.refill<false>();
          if (r == 255u)
            { r = in.hi >> 16; last = in.cons;
              in.take(16);
              in.refill<false>();
            }
          if (r > L - cnt) { r = L - cnt; ln.bad = 1; }
          cnt += r;
          if (wr) out.skip(r,rc4);
          if (cnt >= L) break;
          w = in.hi;
          e = st[w >> 20];
          if (e == 0u) e = long_single6(t2,symtab,w,&ln.bad);
          uint32_t c = e & 0xffu;
          last = in.cons;
          in.take(e >> 8);
          in.refill<false>();
          if (esc && c == 255u)
            { c = in.hi >> 24; last = in.cons;
              in.take(8);
              in.refill<false>();
            }
          if (wr) out.put(c);
          cnt++;
          kept += (c != rc);
        }
      in.top_up();
      if (in.reached() > ln.lim) { ln.bad = 1; break; }
    }
  if (wr) out.flush<true>(rc4);
  *kept_out = kept;
  return (L == 0 || ln.bad) ? ((L == 0) ? 0u : ((last + 47u) >> 5) * 4u) : ((last + 47u) >> 5) * 4u;
}

// the tag line (Unpack_Tag, QV.c:837-847): 'n' where the deletion QV is the run character, the next
// packed tag elsewhere.  The del line is read back from global memory (the lane's own stores) in
// aligned 16-byte blocks; the packed tags are a bit stream of their own.
__device__ __forceinline__ void tag_line(Lane &ln, const uint8_t *del, uint32_t L, int32_t delchar, uint32_t upper)
{ BitIn &in = ln.in; OutSt &out = ln.out;
  const uint32_t caseoff = upper ? 32u : 0u;
  const uint32_t nch = 'n' - caseoff, n4 = nch * 0x01010101u;
  out.prefill(nch);
  const uintptr_t a = reinterpret_cast<uintptr_t>(del);
  const uint4 *blk = reinterpret_cast<const uint4 *>(a & ~(uintptr_t) 15);
  int32_t p = -(int32_t) (a & 15);                    // line position of the block's first byte
  uint32_t polled = 0;
  while (p < (int32_t) L)
    { const int lo = max(0,-p), hi = min(16,(int32_t) L - p);
      uint32_t m = dx_range16(lo,hi);
      if (delchar >= 0)
        { const uint4 v = __ldcg(blk);
          m &= ~dx_eq_mask16(v,(uint32_t) delchar);
        }
      int at = lo;                                    // next block byte not yet written
      while (m)
        { const int k = __ffs(m) - 1; m &= m - 1;
          out.skip((uint32_t) (k - at),n4);
          const uint32_t code = in.hi >> 30;
          in.take(2);
          in.refill<true>();
          out.put(((0x74676361u >> (8u*code)) & 0xffu) - caseoff);
          at = k + 1;
          if (((++polled) & 7u) == 0u) in.top_up();
        }
      out.skip((uint32_t) (hi - at),n4);
      in.top_up();
      blk++; p += 16;
    }
  out.flush<true>(n4);
}

__device__ int fmt_int6(uint8_t *p, int32_t v)
{ char tmp[12];
  int  k = 0, len = 0;
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  if (v < 0) p[len++] = '-';
  do { tmp[k++] = (char) ('0' + u % 10); u /= 10; } while (u);
  while (k) p[len++] = (uint8_t) tmp[--k];
  return len;
}

__global__ void __launch_bounds__(kThreads6,2)
k_qv_decode6(Dec6Args a)
{ Shared6 &sm = *reinterpret_cast<Shared6 *>(dx_dec6_smem);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

  // resident tables: slot 0 del, 1 ins, 2 mrg, 3 sub
  { const QvDecTables4 *T = a.tab;
    for (int s = 0; s < 4; s++)
      { const int symtab = (s == 0) ? 0 : (s == 1) ? 2 : (s == 2) ? 3 : 4;
        const int runtab = (s == 0) ? 1 : 5;
        const bool run = (s == 0 && a.delchar >= 0) || (s == 3 && a.subchar >= 0);
        if (run)
          { const uint32_t *gr = reinterpret_cast<const uint32_t *>(T->single[runtab]);
            const uint32_t *gs = reinterpret_cast<const uint32_t *>(T->single[symtab]);
            for (int j = threadIdx.x; j < 2048; j += kThreads6)
              { sm.tab[s][j] = __ldg(gr + j); sm.tab[s][2048 + j] = __ldg(gs + j); }
          }
        else
          for (int j = threadIdx.x; j < 4096; j += kThreads6) sm.tab[s][j] = __ldg(T->multi[symtab] + j);
      }
  }
  __syncthreads();

  const int64_t t = a.first + ((int64_t) blockIdx.x*kWarps6 + warp)*32 + lane;
  if (t >= a.count) return;
  const int64_t e = (a.order != NULL) ? (int64_t) a.order[t] : t;
  const int32_t Ls = a.rlen[e];
  if (Ls < 0)                                               // ruled out by the host
    { if (a.write != 1) a.status[e] = 1;
      return;
    }
  const uint32_t L = (uint32_t) Ls;
  const QvDecTables2 *t2 = &a.tab->t2;
  const uint8_t *image_end = a.in + a.n;
  Lane ln;
  ln.bad = 0;
  ln.lim = reinterpret_cast<uint64_t>(a.in) + (uint64_t) ((a.limit != NULL) ? a.limit[e] : a.n) + 80u;
  uint32_t *ring_lane = &sm.ring[warp][0][lane];
  uint8_t  *line;
  if (a.write == 2 && a.ent == NULL)
    line = a.out + a.toff[e];
  else
    { const QvDecEntry en = a.ent[e];
      line = a.out + en.text_off;
      if (a.write == 1)
        { uint8_t *h = a.out + en.out_off;                  // "%s/%d/%d_%d RQ=0.%d\n" (undexqv.c:182)
          int hl = 0;
          for (int k = 0; k < a.plen; k++) h[hl++] = (uint8_t) a.prefix[k];
          h[hl++] = '/'; hl += fmt_int6(h+hl,en.well);
          h[hl++] = '/'; hl += fmt_int6(h+hl,en.beg);
          h[hl++] = '_'; hl += fmt_int6(h+hl,en.end);
          const char *rq = " RQ=0.";
          for (int k = 0; k < 6; k++) h[hl++] = (uint8_t) rq[k];
          hl += fmt_int6(h+hl,en.qv);
          h[hl++] = '\n';
        }
    }
  ln.out.init(line,reinterpret_cast<uint8_t *>(sm.outr[warp]) + lane*kOutPitch);

  int64_t at = a.start[e];
  int64_t o[6];
  for (int k = 0; k < 6; k++) o[k] = at;
  const uint16_t *tab16[4];
  for (int s = 0; s < 4; s++) tab16[s] = reinterpret_cast<const uint16_t *>(sm.tab[s]);

  do
    { uint32_t kept = L, bytes;
      // deletion QVs
      if (at > a.n) { ln.bad = 1; break; }
      if (L > 0) ln.in.init<false>(a.in + at,image_end,ring_lane);
      if (a.delchar >= 0)
        bytes = run_stream(ln,tab16[0],tab16[0] + 4096,t2,0,1,t2->type[0] == 2,(uint32_t) a.delchar,L,true,&kept);
      else
        bytes = plain_stream(ln,sm.tab[0],t2,0,L,true,NULL);
      ln.out.put('\n');
      at += bytes; o[1] = at;
      if (ln.bad && a.write != 1) break;
      // deletion tags
      const uint32_t clen = (a.delchar < 0) ? L : kept;
      const int64_t tbytes = (clen + 3) >> 2;
      if (at + tbytes <= a.n)
        { if (L > 0)
            { ln.out.sync_partial();                         // the del line is complete in global memory
              ln.in.init<true>(a.in + at,image_end,ring_lane);
              tag_line(ln,line,L,a.delchar,(uint32_t) a.upper);
            }
        }
      else
        { ln.bad = 1;
          for (uint32_t k = 0; k < L; k++) ln.out.put('n'), ln.out.flush<false>(0);
        }
      ln.out.put('\n');
      at += tbytes; o[2] = at;
      if (at > a.n) { ln.bad = 1; if (a.write != 1) break; at = a.n; }
      // insertion and merge QVs
      if (L > 0) ln.in.init<false>(a.in + at,image_end,ring_lane);
      bytes = plain_stream(ln,sm.tab[1],t2,2,L,true,NULL);
      ln.out.put('\n');
      at += bytes; o[3] = at;
      if (ln.bad && a.write != 1) break;
      if (at > a.n) { ln.bad = 1; if (a.write != 1) break; at = a.n; }
      if (L > 0) ln.in.init<false>(a.in + at,image_end,ring_lane);
      bytes = plain_stream(ln,sm.tab[2],t2,3,L,true,NULL);
      ln.out.put('\n');
      at += bytes; o[4] = at;
      if (ln.bad && a.write != 1) break;
      if (at > a.n) { ln.bad = 1; if (a.write != 1) break; at = a.n; }
      // substitution QVs
      if (L > 0) ln.in.init<false>(a.in + at,image_end,ring_lane);
      if (a.subchar >= 0)
        bytes = run_stream(ln,tab16[3],tab16[3] + 4096,t2,4,5,t2->type[4] == 2,(uint32_t) a.subchar,L,true,&kept);
      else
        bytes = plain_stream(ln,sm.tab[3],t2,4,L,true,NULL);
      ln.out.put('\n');
      at += bytes; o[5] = at;
      ln.out.sync_partial();
    }
  while (false);

  const int64_t lim = (a.limit != NULL) ? a.limit[e] : a.n;
  if (at > lim) ln.bad = 1;
  if (a.soff != NULL)
    for (int k = 0; k < 6; k++) a.soff[e*6 + k] = (k == 0 || o[k] >= o[k-1]) ? o[k] : at;
  if (a.write == 1) { if (ln.bad) atomicExch(a.status,1); }
  else a.status[e] = (int32_t) ln.bad;
}

}  // namespace

// Tickets [0, n_coop) -- the longest entries -- go to the warp-per-entry kernel, the rest to the
// lane-per-entry kernel.  Same contract as dxk_qv_decode5x otherwise (write = 1 or 2).
int dxk_qv_decode6x(dx_ctx *ctx, const uint8_t *d_in, size_t n, const QvDecTables4 *d_tab,
                    int delchar, int subchar, int upper, int write, int64_t count,
                    const int64_t *d_start, const int32_t *d_rlen, const QvDecEntry *d_ent,
                    const char *d_prefix, int plen, uint8_t *d_out, int64_t *d_soff, int32_t *d_status,
                    const int64_t *d_limit, const int32_t *d_order, const int64_t *d_toff, int64_t n_coop)
{ if (count == 0) return DX_OK;
  int rc;
  if (n_coop > count) n_coop = count;
  if (n_coop > 0 &&
      (rc = dxk_qv_decode5x(ctx,d_in,n,d_tab,delchar,subchar,upper,write,n_coop,d_start,d_rlen,d_ent,d_prefix,plen,
                            d_out,d_soff,d_status,d_limit,d_order,d_toff)) != DX_OK) return rc;
  if (n_coop >= count) return DX_OK;
  Dec6Args a;
  a.in = d_in; a.n = (int64_t) n; a.tab = d_tab;
  a.delchar = delchar; a.subchar = subchar; a.upper = upper; a.write = write;
  a.first = n_coop; a.count = count; a.start = d_start; a.rlen = d_rlen; a.ent = d_ent;
  a.toff = d_toff; a.prefix = d_prefix; a.plen = plen; a.out = d_out; a.soff = d_soff; a.status = d_status;
  a.limit = d_limit; a.order = d_order;
  const size_t smem = sizeof(Shared6);
  DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_decode6,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem));
  const int64_t grid = (count - n_coop + kThreads6 - 1) / kThreads6;
  DX_PROF_BEGIN(ctx); k_qv_decode6<<<(unsigned) grid,kThreads6,smem,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,write == 1 ? "k_qv_decode6" : "k_qv_decode6_spec");
  return DX_OK;
}
