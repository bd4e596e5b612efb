// dx_qv_encode.cu -- pass 2 of the QV coder on the device.
//
// Replaces Encode / Encode_Run (reference QV.c:386-506), Pack_Tag + Number_Read + Compress_Read
// on the tag line (QV.c:810-819, 1400-1404), the lossy masks (QV.c:1406-1415) and the per-entry
// header writes of dexqv.c:128-139.
//
//   k_qv_size     one warp per (entry, stream): bits of every item, position of the last item
//                 -> 32-bit words the reference would write (the (p_last+47)>>5 rule,
//                 QV.c:436-442), kept-tag count -> tag bytes.  No bit positions are needed here,
//                 so lanes only sum code lengths.
//   k_qv_offsets  exclusive scan of the per-entry byte totals (the implicit file position of
//                 the reference's fwrite calls)
//   k_qv_emit     one warp per (entry, stream).  Plain streams: every lane looks up the codes of
//                 16 consecutive symbols, a warp scan of the bit counts gives each lane its bit
//                 offset, the lane shifts its codes (two per step) through a 64-bit register and
//                 stores the words it COMPLETES into the warp's staging area; the partial words
//                 between neighbouring lanes are merged by a segmented OR-scan over shuffles, so
//                 no shared-memory atomics are involved.  Run-length streams first compact the
//                 (position, symbol) pairs that are not the run character into a per-warp queue,
//                 then code 32 items at a time, one item (run code [+ literal] + symbol code) per
//                 lane.  The stage is flushed with aligned 32-bit stores at whatever byte alignment
//                 the stream has in the file.
//
// Two routes.  The usual one needs no size pass: k_qv_emit codes every stream into a scratch image
// laid out like the text (a stream of rlen symbols has rlen+9 bytes of room, i.e. more than 8 bits per
// symbol) and records its length, k_qv_offsets scans the lengths and k_qv_compact moves the streams
// to their place in the file with 16-byte stores.  A stream that does not fit (possible only when an
// entry is full of escaped symbols) sends the call to the exact route: k_qv_size, k_qv_offsets,
// k_qv_emit straight into the file.

#include "dx_internal.h"
#include "dx_common.cuh"
#include "dx_bits.cuh"

namespace {

constexpr int kEncWarps   = 16;
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kStageWords = 512;                  // per warp; a 512-symbol row needs <= 384
constexpr int kQueue      = 1024;                 // per warp ring of (position << 8 | symbol)
constexpr int kWarpWords  = kStageWords + 4 + kQueue;

struct EncArgs
{ const uint8_t  *text;
  const uint8_t  *text_end16;     // first address past the readable 16-byte-padded text
  QvEntries       ent;
  const uint32_t *tab;            // [6][256] packed entries (DX_ENC2_*)
  int32_t         delchar, subchar, lossy, lwell_in;
  uint32_t       *bytes;          // [n][6]
  int64_t        *off;            // [n+1]
  uint8_t        *out;            // MODE 1: the file image; MODE 2: the scratch image
  int32_t        *ovf;            // MODE 2: set when a stream did not fit its room in the scratch image
  unsigned long long *ticket;
};

// MODE 2: where the stream of line s (0 del 1 tag 2 ins 3 mrg 4 sub) of entry e goes in the scratch
// image, and how many bytes it may take
__device__ __forceinline__ int64_t scratch_off(const EncArgs &a, int64_t e, int s, int32_t rlen)
{ return a.ent.line0[e] + (int64_t) s*((int64_t) rlen + 1) + 8*(5*e + s); }
__device__ __forceinline__ uint32_t scratch_room(int32_t rlen) { return (uint32_t) rlen + 9u; }

// table entry: bits 0-4 length of the whole item, bit 5 escape, bits 8-31 the item's bits.
// Symbol tables fold the 8-bit literal of an escaped symbol into the item (length <= 24); run
// tables keep only the code, the 16-bit literal run length follows as a second piece.
#define E_LEN(e)   ((e) & 31u)
#define E_ESC(e)   ((e) & 32u)
#define E_BITS(e)  ((e) >> 8)

__device__ __forceinline__ uint32_t base2(uint32_t c)        // Number_Read, DB.c:394-411
{ c |= 0x20u;
  return (c == 'c') ? 1u : (c == 'g') ? 2u : (c == 't') ? 3u : 0u;
}

// ---- plain stream: one Huffman item per symbol (Encode, QV.c:386-443) --------------------------------
// One row (32 lanes x 16 bytes) per iteration, the next row's chunk already in flight.  The code is
// kept small on purpose: the kernel is instruction-cache bound when every loop is unrolled.
template <int MODE>
__device__ void code_plain(const EncArgs &a, const uint32_t *sym, const uint8_t *line, int32_t rlen,
                           uint32_t lossmask, int lane, WarpBits &wb, uint32_t &total_bits,
                           uint32_t &plast_out)
{ LineWalk lw; lw.set(line,rlen);
  uint32_t mybits = 0;                                   // MODE 0: lane-local sum over the line
  uint4 nxt = (lane < lw.nchunk) ? dx_ldg16(lw.base + (int64_t) lane*16) : make_uint4(0,0,0,0);
#pragma unroll 1
  for (int32_t c0 = 0; c0 < lw.nchunk; c0 += 32)
    { const int32_t c = c0 + lane;
      const uint4 v = nxt;
      nxt = (c + 32 < lw.nchunk) ? dx_ldg16(lw.base + (int64_t) (c + 32)*16) : make_uint4(0,0,0,0);
      const uint32_t valid = lw.valid(c);
      uint32_t w[4] = { v.x & lossmask, v.y & lossmask, v.z & lossmask, v.w & lossmask };
      if (valid != 0xffffu)                                // first / last chunk of the line (or none of it):
        {                                                  // byte 0 has no code (see dxk_qv_encode), so blank
#pragma unroll
          for (int q = 0; q < 4; q++)
            { const uint32_t n4 = (valid >> (4*q)) & 15u;
              w[q] &= ((n4 & 1u) ? 0xffu : 0u) | ((n4 & 2u) ? 0xff00u : 0u) |
                      ((n4 & 4u) ? 0xff0000u : 0u) | ((n4 & 8u) ? 0xff000000u : 0u);
            }
        }
      uint32_t e[16];
      uint32_t bits = 0, esc = 0;
#pragma unroll
      for (int i = 0; i < 16; i++)
        { const uint32_t t = sym[__byte_perm(w[i >> 2],0u,0x4440 + (i & 3))];
          e[i] = t;
          bits += E_LEN(t);
          esc  |= t;
        }
      if (MODE == 0) { mybits += bits; continue; }
      const uint32_t inc = dx_warp_incl_sum(bits,lane);
      const uint32_t row = __shfl_sync(DX_FULL,inc,31);
      wb.reserve<false>(row,lane);
      LaneSink sk;
      sk.start(wb.bitpos() + inc - bits);
      if (!__any_sync(DX_FULL,E_ESC(esc) != 0u))
        {
#pragma unroll
          for (int i = 0; i < 16; i += 2)                  // codes <= 16 bits: two per step
            { const uint32_t l1 = E_LEN(e[i+1]);
              sk.put(wb.stage,(E_BITS(e[i]) << l1) | E_BITS(e[i+1]),E_LEN(e[i]) + l1);
            }
        }
      else                                                 // a literal somewhere in the row (rare)
        {
#pragma unroll 1
          for (int i = 0; i < 16; i++)
            { uint32_t t = sym[dx_byte_of(v,i) & lossmask & 0xffu];
              if (!((valid >> i) & 1u)) t = 0;
              sk.put(wb.stage,E_BITS(t),E_LEN(t));
            }
        }
      if (__all_sync(DX_FULL,bits >= 32u)) sk.finish_wide(wb,lane);
      else                                 sk.finish(wb,lane);
    }
  if (MODE == 0) total_bits = dx_warp_sum(mybits);
  else           total_bits = wb.total();
  // the last item: the last symbol's code, or its literal when it is escaped
  const uint32_t el = sym[line[rlen-1] & lossmask & 0xffu];
  plast_out = total_bits - E_LEN(el) + (E_ESC(el) ? E_LEN(el) - 8u : 0u);
}

// ---- run-length stream (Encode_Run, QV.c:448-506) -----------------------------------------------------
// items = (run of the run character, next other symbol); a trailing run has no symbol
template <int MODE>
__device__ void code_run(const EncArgs &a, const uint32_t *sym, const uint32_t *run, uint32_t rc,
                         const uint8_t *line, int32_t rlen, int lane, uint32_t *queue, WarpBits &wb,
                         uint32_t &total_bits, uint32_t &plast_out)
{ LineWalk lw; lw.set(line,rlen);
  uint32_t qhead = 0, qtail = 0;                         // ring positions (warp-uniform)
  int32_t  prevpos = -1;                                 // last position that was not the run character
  uint32_t mybits = 0;
  uint4 nxt = (lane < lw.nchunk) ? dx_ldg16(lw.base + (int64_t) lane*16) : make_uint4(0,0,0,0);
#pragma unroll 1
  for (int32_t c0 = 0; c0 < lw.nchunk + 32; c0 += 32)    // one extra round drains the queue
    { const bool last = (c0 >= lw.nchunk);
      if (!last)
        { const int32_t c = c0 + lane;
          const uint4 v = nxt;
          nxt = (c + 32 < lw.nchunk) ? dx_ldg16(lw.base + (int64_t) (c + 32)*16) : make_uint4(0,0,0,0);
          const int32_t p0 = c*16 - lw.skew;
          uint32_t m = lw.valid(c) & ~dx_eq_mask16(v,rc);
          const uint32_t cnt = __popc(m);
          const uint32_t inc = dx_warp_incl_sum(cnt,lane);
          uint32_t at = qtail + inc - cnt;
          while (m)
            { const int i = __ffs(m) - 1; m &= m - 1;
              queue[at & (kQueue-1)] = ((uint32_t) (p0 + i) << 8) | dx_byte_of(v,i);
              at++;
            }
          qtail += __shfl_sync(DX_FULL,inc,31);
          __syncwarp();
        }
      // code the queued items 32 at a time, lane i takes item qhead + i
#pragma unroll 1
      while (qtail - qhead >= 32u || (last && qtail != qhead))
        { const uint32_t n = min(32u,qtail - qhead);
          uint32_t R = 0, E = 0, r = 0, bits = 0;
          int32_t p = 0;
          if ((uint32_t) lane < n)
            { const uint32_t it = queue[(qhead + lane) & (kQueue-1)];
              p = (int32_t) (it >> 8);
              const int32_t pp = (lane == 0) ? prevpos : (int32_t) (queue[(qhead + lane - 1) & (kQueue-1)] >> 8);
              r = (uint32_t) (p - pp - 1);
              R = run[min(r,255u)];
              E = sym[it & 0xffu];
              bits = E_LEN(R) + (E_ESC(R) ? 16u : 0u) + E_LEN(E);
            }
          prevpos = __shfl_sync(DX_FULL,p,n-1);
          qhead += n;
          if (MODE == 0) { mybits += bits; continue; }
          const uint32_t inc = dx_warp_incl_sum(bits,lane);
          wb.reserve<false>(__shfl_sync(DX_FULL,inc,31),lane);
          LaneSink sk;
          sk.start(wb.bitpos() + inc - bits);
          sk.put(wb.stage,E_BITS(R) & 0xffffu,E_LEN(R));
          if (E_ESC(R)) sk.put(wb.stage,r & 0xffffu,16);
          sk.put(wb.stage,E_BITS(E),E_LEN(E));
          sk.finish(wb,lane);
        }
    }

  // trailing run of the run character (QV.c:475-487 when k reaches rlen inside a run)
  uint32_t Rt = 0, rt = 0;
  const bool trailing = (prevpos < rlen-1);
  if (trailing)
    { rt = (uint32_t) (rlen-1-prevpos);
      Rt = run[min(rt,255u)];
      const uint32_t tb = E_LEN(Rt) + (E_ESC(Rt) ? 16u : 0u);
      if (MODE == 0) mybits += (lane == 0) ? tb : 0u;
      else
        { wb.reserve<false>(32u,lane);
          LaneSink sk;
          sk.start(wb.bitpos() + (lane == 0 ? 0u : tb));
          if (lane == 0)
            { sk.put(wb.stage,E_BITS(Rt) & 0xffffu,E_LEN(Rt));
              if (E_ESC(Rt)) sk.put(wb.stage,rt & 0xffffu,16);
            }
          sk.finish(wb,lane);
        }
    }
  if (MODE == 0) total_bits = dx_warp_sum(mybits);
  else           total_bits = wb.total();
  if (trailing)
    plast_out = total_bits - (E_ESC(Rt) ? 16u : E_LEN(Rt));
  else
    { const uint32_t el = sym[line[rlen-1]];
      plast_out = total_bits - E_LEN(el) + (E_ESC(el) ? E_LEN(el) - 8u : 0u);
    }
}

template <int MODE>
__device__ void code_stream(const EncArgs &a, const uint32_t *stab, const uint8_t *line,
                            int32_t rlen, int kind /*0 del 2 ins 3 mrg 4 sub*/, int lane,
                            uint32_t *stage, uint32_t *queue, uint8_t *gptr,
                            uint32_t &total_bits, uint32_t &plast_out, uint32_t &ovf)
{ total_bits = 0; plast_out = 0; ovf = 0;
  if (rlen <= 0) return;
  const int32_t rc = (kind == 0) ? a.delchar : (kind == 4) ? a.subchar : -1;
  const uint32_t lossmask = !a.lossy ? 0xffffffffu : (kind == 2) ? 0xfefefefeu
                                                    : (kind == 3) ? 0xfcfcfcfcu : 0xffffffffu;
  WarpBits wb; wb.init(stage,gptr,kStageWords);
  if (MODE == 2) wb.limit = scratch_room(rlen);
  if (rc < 0) code_plain<MODE>(a,stab + kind*256,line,rlen,lossmask,lane,wb,total_bits,plast_out);
  else        code_run<MODE>(a,stab + kind*256,stab + (kind == 0 ? 1 : 5)*256,(uint32_t) rc,line,rlen,lane,
                             queue,wb,total_bits,plast_out);
  if (MODE >= 1)
    { // final flush incl. the look-ahead padding word (QV.c:436-442)
      __syncwarp();
      const uint32_t full_total = (total_bits + 31u) >> 5;
      const uint32_t want_total = (plast_out + 47u) >> 5;
      if (lane == 0)
        { if (wb.cbits) stage[wb.nst] = wb.carry;
          if (want_total > full_total) stage[wb.nst + (wb.cbits ? 1u : 0u)] = wb.cbits ? wb.carry : 0u;
        }
      uint32_t nst = wb.nst + (wb.cbits ? 1u : 0u) + (want_total > full_total ? 1u : 0u);
      __syncwarp();
      if ((wb.flushed + nst)*4u <= wb.limit) copy_out<false>(gptr + (size_t) wb.flushed*4u,stage,nst*4u,lane);
      else                                   wb.ovf = 1;
      __syncwarp();
      ovf = wb.ovf;
    }
}

__device__ __forceinline__ uint32_t stream_words(uint32_t total_bits, uint32_t plast, int32_t rlen)
{ if (rlen <= 0) return 0;
  const uint32_t full = (total_bits + 31u) >> 5, want = (plast + 47u) >> 5;
  return max(full,want);
}

// kept tags of one entry: count (MODE 0) or 2-bit pack into global (MODE 1)
template <int MODE>
__device__ uint32_t code_tags(const EncArgs &a, const uint8_t *del, const uint8_t *tag,
                              int32_t rlen, int lane, uint32_t *stage, uint8_t *gptr)
{ if (MODE == 0 && a.delchar < 0) return (uint32_t) rlen;      // every tag is kept
  LineWalk lw; lw.set(tag,rlen);
  WarpBits wb; wb.init(stage,gptr,kStageWords);
  uint32_t mykept = 0;
#pragma unroll 1
  for (int32_t c0 = 0; c0 < lw.nchunk; c0 += 32)
    { const int32_t c = c0 + lane;
      const int32_t p0 = c*16 - lw.skew;
      uint32_t m = lw.valid(c);
      uint4 tv = make_uint4(0,0,0,0);
      if (c < lw.nchunk)
        { if (MODE == 1) tv = dx_ldg16(lw.base + (int64_t) c*16);
          if (a.delchar >= 0)
            { uint4 dv = dx_ld16_any(del + p0,a.text_end16);    // del bytes at the same positions
              m &= ~dx_eq_mask16(dv,(uint32_t) a.delchar);
            }
        }
      const uint32_t cnt = __popc(m);
      if (MODE == 0) { mykept += cnt; continue; }
      const uint32_t inc = dx_warp_incl_sum(cnt,lane);
      const uint32_t row = __shfl_sync(DX_FULL,inc,31);
      wb.reserve<true>(row*2u,lane);
      uint32_t val = 0;
      while (m)
        { const int i = __ffs(m) - 1; m &= m - 1;
          val = (val << 2) | base2(dx_byte_of(tv,i));
        }
      LaneSink sk;
      sk.start(wb.bitpos() + (inc - cnt)*2u);
      sk.put(wb.stage,val,cnt*2u);
      sk.finish(wb,lane);
    }
  if (MODE == 0) return dx_warp_sum(mykept);
  __syncwarp();
  const uint32_t kept = wb.total() >> 1;
  if (lane == 0 && wb.cbits) stage[wb.nst] = wb.carry;
  __syncwarp();
  // MSB-first inside BYTES: the staged big-endian words go out byte-swapped; only the bytes in use
  const uint32_t nbytes = ((kept + 3u) >> 2) - wb.flushed*4u;
  copy_out<true>(gptr + (size_t) wb.flushed*4u,stage,nbytes,lane);
  __syncwarp();
  return kept;
}

__device__ __forceinline__ uint32_t well_bytes(const EncArgs &a, int64_t e)
{ const int32_t lw = (e == 0) ? a.lwell_in : a.ent.well[e-1];
  const int32_t d  = a.ent.well[e] - lw;
  return 1u + (d >= 255 ? (uint32_t) d / 255u : 0u);            // dexqv.c:128-135
}

template <int MODE>
__global__ void __launch_bounds__(kEncThreads,2)
k_qv_code(EncArgs a)
{ extern __shared__ uint32_t smem[];
  uint32_t *stab  = smem;                                          // [6][256]
  uint32_t *stage = smem + 6*256 + (threadIdx.x >> 5)*kWarpWords;  // per warp: stage, then the item queue
  uint32_t *queue = stage + kStageWords + 4;
  for (int i = threadIdx.x; i < 6*256; i += kEncThreads) stab[i] = a.tab[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t nunits = a.ent.n * 5;
  unsigned long long next = 0;
  if (lane == 0) next = atomicAdd(a.ticket,1ull);
  while (true)
    { const int64_t u = (int64_t) __shfl_sync(DX_FULL,next,0);
      if (u >= nunits) break;
      if (lane == 0) next = atomicAdd(a.ticket,1ull);              // in flight while this unit is coded
      const int64_t t = u / 5;
      const int     s = (int) (u - t*5);                           // 0 del 1 tag 2 ins 3 mrg 4 sub
      const int64_t e = (a.ent.order != NULL) ? (int64_t) a.ent.order[t] : t;
      const int32_t rlen = a.ent.rlen[e];
      const uint8_t *l0  = a.text + a.ent.line0[e];
      const uint8_t *line = l0 + (int64_t) s*((int64_t) rlen + 1);
      uint8_t *gptr = NULL;
      if (MODE == 1)
        { int64_t o = a.off[e];
          for (int k = 0; k <= s; k++) o += a.bytes[e*6 + k];
          gptr = a.out + o;
        }
      if (MODE == 2) gptr = a.out + scratch_off(a,e,s,rlen);
      if (s == 1)
        { uint32_t kept = code_tags<(MODE == 0) ? 0 : 1>(a,l0,line,rlen,lane,stage,gptr);   // (rlen+3)/4 bytes at most
          if (MODE != 1 && lane == 0) a.bytes[e*6 + 2] = (kept + 3u) >> 2;
        }
      else
        { uint32_t bits, plast, ovf;
          code_stream<MODE>(a,stab,line,rlen,s,lane,stage,queue,gptr,bits,plast,ovf);
          if (MODE != 1 && lane == 0)
            { a.bytes[e*6 + 1 + s] = stream_words(bits,plast,rlen)*4u;
              if (s == 0) a.bytes[e*6] = well_bytes(a,e) + 12u;
              if (MODE == 2 && ovf) atomicExch(a.ovf,1);
            }
          if (MODE == 1 && s == 0 && lane == 0)
            { // entry header: well-delta bytes, beg, end, qv (dexqv.c:128-139)
              uint8_t *h = a.out + a.off[e];
              int32_t lwell = (e == 0) ? a.lwell_in : a.ent.well[e-1];
              const int32_t well = a.ent.well[e];
              while (well - lwell >= 255) { *h++ = 0xff; lwell += 255; }
              *h++ = (uint8_t) (well - lwell);
              const int32_t f3[3] = { a.ent.beg[e], a.ent.end[e], a.ent.qv[e] };
              for (int k = 0; k < 3; k++)
                for (int b = 0; b < 4; b++)
                  *h++ = (uint8_t) ((uint32_t) f3[k] >> (8*b));
            }
        }
      __syncwarp();
    }
}

// exclusive scan of per-entry byte totals; one CTA
__global__ void __launch_bounds__(1024)
k_qv_offsets(const uint32_t *bytes, int64_t n, int64_t *off)
{ __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < n; b += 1024)
    { const int64_t i = b + threadIdx.x;
      uint64_t v = 0;
      if (i < n)
        for (int k = 0; k < 6; k++) v += bytes[i*6 + k];
      const uint64_t inc = dx_warp_incl_sum64(v,lane);
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      if (warp == 0)
        { const uint64_t w = wsum[lane];
          const uint64_t wi = dx_warp_incl_sum64(w,lane);
          wsum[lane] = wi - w;
        }
      __syncthreads();
      const uint64_t excl = carry + wsum[warp] + inc - v;
      if (i < n) off[i] = (int64_t) excl;
      __syncthreads();
      if (threadIdx.x == 1023) carry = excl + v;
      __syncthreads();
    }
  if (threadIdx.x == 0) off[n] = (int64_t) carry;
}

// per-entry byte totals (header + 5 streams) for the multi-CTA scan of dx_frame.cu
__global__ void k_qv_entry_total(const uint32_t *bytes, int64_t n, uint32_t *total)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t v = 0;
  for (int k = 0; k < 6; k++) v += bytes[i*6 + k];
  total[i] = v;
}

// ---- scratch image -> file image -----------------------------------------------------------------------
// one warp per (entry, line): the stream k_qv_code<2> left in the scratch image moves to its place
// behind the entry's header (16-byte stores, source at any alignment); line 0 also writes the header
struct CompactArgs
{ const uint8_t  *scratch;
  const uint8_t  *scratch_end16;
  QvEntries       ent;
  const uint32_t *bytes;          // [n][6]
  const int64_t  *off;            // [n+1]
  int32_t         lwell_in;
  uint8_t        *out;
  unsigned long long *ticket;
  const int32_t  *ovf;            // a stream outgrew its scratch room: nothing to move, the host takes the exact route
  int64_t         cap;            // room at out: a larger image is an error the host reports, nothing is written
};

__global__ void __launch_bounds__(256)
k_qv_compact(CompactArgs a)
{ const int lane = threadIdx.x & 31;
  // launched before the host has seen the total (one round trip less): the same two tests here
  if (*a.ovf != 0 || a.off[a.ent.n] > a.cap) return;
  const int64_t nunits = a.ent.n * 5;
  unsigned long long next = 0;
  if (lane == 0) next = atomicAdd(a.ticket,1ull);
  while (true)
    { const int64_t u = (int64_t) __shfl_sync(DX_FULL,next,0);
      if (u >= nunits) break;
      if (lane == 0) next = atomicAdd(a.ticket,1ull);
      const int64_t t = u / 5;
      const int     s = (int) (u - t*5);
      const int64_t e = (a.ent.order != NULL) ? (int64_t) a.ent.order[t] : t;
      const int32_t rlen = a.ent.rlen[e];
      int64_t o = a.off[e];
      for (int k = 0; k <= s; k++) o += a.bytes[e*6 + k];
      const uint32_t n = a.bytes[e*6 + 1 + s];
      if (s == 0 && lane == 0)
        { // entry header: well-delta bytes, beg, end, qv (dexqv.c:128-139)
          uint8_t *h = a.out + a.off[e];
          int32_t lwell = (e == 0) ? a.lwell_in : a.ent.well[e-1];
          const int32_t well = a.ent.well[e];
          while (well - lwell >= 255) { *h++ = 0xff; lwell += 255; }
          *h++ = (uint8_t) (well - lwell);
          const int32_t f3[3] = { a.ent.beg[e], a.ent.end[e], a.ent.qv[e] };
          for (int k = 0; k < 3; k++)
            for (int b = 0; b < 4; b++)
              *h++ = (uint8_t) ((uint32_t) f3[k] >> (8*b));
        }
      if (n == 0) continue;
      const uint8_t *src = a.scratch + (a.ent.line0[e] + (int64_t) s*((int64_t) rlen + 1) + 8*(5*e + s));
      uint8_t *dst = a.out + o;
      uint32_t head = (16u - (uint32_t) (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
      if (head > n) head = n;
      if ((uint32_t) lane < head) dst[lane] = src[lane];
      const uint32_t nvec = (n - head) >> 4;
      uint32_t i = lane;
      for ( ; i + 96 < nvec; i += 128)                      // four independent 16-byte moves in flight
        { const uint4 v0 = dx_ld16_any(src + head + (size_t) i*16,a.scratch_end16),
                      v1 = dx_ld16_any(src + head + (size_t) (i+32)*16,a.scratch_end16),
                      v2 = dx_ld16_any(src + head + (size_t) (i+64)*16,a.scratch_end16),
                      v3 = dx_ld16_any(src + head + (size_t) (i+96)*16,a.scratch_end16);
          dx_stg16(dst + head + (size_t) i*16,v0);      dx_stg16(dst + head + (size_t) (i+32)*16,v1);
          dx_stg16(dst + head + (size_t) (i+64)*16,v2); dx_stg16(dst + head + (size_t) (i+96)*16,v3);
        }
      for ( ; i < nvec; i += 32)
        dx_stg16(dst + head + (size_t) i*16,dx_ld16_any(src + head + (size_t) i*16,a.scratch_end16));
      const uint32_t done = head + nvec*16u;
      if ((uint32_t) lane < n - done) dst[done + lane] = src[done + lane];
    }
}

}  // namespace

int dxk_qv_encode(dx_ctx *ctx, const uint8_t *d_text, size_t text_n, QvEntries ent,
                  const QvEncTables *h_tab, int delchar, int subchar, int lossy, int32_t lwell_in,
                  uint8_t *d_out, size_t cap, size_t *out_len, int32_t *last_well,
                  int64_t *h_entry_off, int64_t max_entries)
{ *out_len = 0;
  if (last_well) *last_well = lwell_in;
  if (ent.n == 0)
    { if (h_entry_off && max_entries >= 0) h_entry_off[0] = 0;
      return DX_OK;
    }
  const int64_t n = ent.n;
  uint32_t *d_tab   = (uint32_t *) dx_arena_get(ctx,sizeof(QvEncTables));
  uint32_t *d_bytes = (uint32_t *) dx_arena_get(ctx,(size_t) n*6*4);
  int64_t  *d_off   = (int64_t *)  dx_arena_get(ctx,(size_t) (n+1)*8);
  unsigned long long *d_ticket = (unsigned long long *) dx_arena_get(ctx,16);
  if (!d_tab || !d_bytes || !d_off || !d_ticket) return DX_E_NOMEM;
  // k_qv_code blanks the bytes of a chunk that lie outside the line and relies on byte 0 having no
  // code; a NUL inside a line would end the reference's fgets string (QV.c:751-798), so no valid
  // coding gives it one
  for (int k = 0; k < 6; k++)
    if (k != 1 && k != 5 && h_tab->t[k][0] != 0)
      return dx_fail(ctx,DX_E_CODING,"the coding scheme assigns a code to the NUL byte");
  DX_CUDA(ctx,cudaMemcpyAsync(d_tab,h_tab,sizeof(QvEncTables),cudaMemcpyHostToDevice,ctx->stream));
  DX_CUDA(ctx,cudaMemsetAsync(d_ticket,0,16,ctx->stream));

  EncArgs a;
  a.text = d_text;
  a.text_end16 = d_text + ((text_n + 15) & ~(size_t) 15);
  a.ent = ent; a.tab = d_tab;
  a.delchar = delchar; a.subchar = subchar; a.lossy = lossy; a.lwell_in = lwell_in;
  a.bytes = d_bytes; a.off = d_off; a.out = d_out; a.ovf = NULL; a.ticket = d_ticket;

  const int grid = ctx->sm_count * 2;
  const size_t smem1 = 6*256*4 + (size_t) kEncWarps*kWarpWords*4;
  const size_t smem0 = smem1;
  DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_code<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem0));
  DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_code<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem1));
  DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_code<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem1));

  int64_t total = 0;
  int32_t lastw = 0;
  bool done = false;
  int rc = DX_OK;
  uint32_t *d_tot = (uint32_t *) dx_arena_get(ctx,(size_t) n*4);
  if (d_tot == NULL) return DX_E_NOMEM;
  auto entry_offsets = [&]() -> int                     // exclusive scan of the entries' byte totals
    { if (n < 16384)                                      // few entries: the one-CTA kernel
        { DX_PROF_BEGIN(ctx); k_qv_offsets<<<1,1024,0,ctx->stream>>>(d_bytes,n,d_off);
          DX_LAUNCHED(ctx,"k_qv_offsets");
          return DX_OK;
        }
      DX_PROF_BEGIN(ctx);
      k_qv_entry_total<<<(unsigned) ((n + 255)/256),256,0,ctx->stream>>>(d_bytes,n,d_tot);
      DX_LAUNCHED(ctx,"k_qv_offsets");
      return dxk_scan_u32(ctx,d_tot,n,d_off);
    };
  if (!ctx->route[DXR_TWO_PASS])
    { // code into a scratch image first, then move the streams to where their lengths put them
      const size_t sbytes = ((text_n + (size_t) n*40 + 15) & ~(size_t) 15) + 32;
      uint8_t *d_scratch = (uint8_t *) dx_arena_get(ctx,sbytes);
      int32_t *d_ovf = (int32_t *) dx_arena_get(ctx,16);
      unsigned long long *d_ticket2 = (unsigned long long *) dx_arena_get(ctx,16);
      if (!d_scratch || !d_ovf || !d_ticket2) return DX_E_NOMEM;
      DX_CUDA(ctx,cudaMemsetAsync(d_ovf,0,16,ctx->stream));
      DX_CUDA(ctx,cudaMemsetAsync(d_ticket2,0,16,ctx->stream));
      EncArgs b = a;
      b.out = d_scratch; b.ovf = d_ovf; b.ticket = d_ticket2;
      DX_PROF_BEGIN(ctx); k_qv_code<2><<<grid,kEncThreads,smem1,ctx->stream>>>(b);
      DX_LAUNCHED(ctx,"k_qv_emit");
      if ((rc = entry_offsets()) != DX_OK) return rc;
      struct Res { int64_t total; int32_t lastw, ovf; };
      Res *hr = (Res *) dx_hpin_get(ctx,sizeof(Res));
      if (hr == NULL) return DX_E_NOMEM;
      CompactArgs c;
      c.scratch = d_scratch; c.scratch_end16 = d_scratch + sbytes - 16;
      c.ent = ent; c.bytes = d_bytes; c.off = d_off; c.lwell_in = lwell_in; c.out = d_out;
      c.ticket = d_ticket2 + 1; c.ovf = d_ovf; c.cap = (int64_t) cap;
      DX_PROF_BEGIN(ctx); k_qv_compact<<<ctx->sm_count*8,256,0,ctx->stream>>>(c);
      DX_LAUNCHED(ctx,"k_qv_compact");
      DX_CUDA(ctx,cudaMemcpyAsync(&hr->total,d_off+n,8,cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaMemcpyAsync(&hr->lastw,ent.well+(n-1),4,cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaMemcpyAsync(&hr->ovf,d_ovf,4,cudaMemcpyDeviceToHost,ctx->stream));
      if (h_entry_off != NULL && max_entries >= n)
        DX_CUDA(ctx,cudaMemcpyAsync(h_entry_off,d_off,(size_t) (n+1)*8,cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
      total = hr->total; lastw = hr->lastw;
      if (!hr->ovf)
        { if ((size_t) total > cap)
            return dx_fail_cap(ctx,(size_t) (total),cap);
          done = true;
          if (h_entry_off != NULL && max_entries >= n)     // everything is on the host already
            { *out_len = (size_t) total;
              if (last_well) *last_well = lastw;
              return DX_OK;
            }
        }
    }
  if (!done)
    { DX_PROF_BEGIN(ctx); k_qv_code<0><<<grid,kEncThreads,smem0,ctx->stream>>>(a);
      DX_LAUNCHED(ctx,"k_qv_size");
      if ((rc = entry_offsets()) != DX_OK) return rc;
      DX_CUDA(ctx,cudaMemcpyAsync(&total,d_off+n,8,cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaMemcpyAsync(&lastw,ent.well+(n-1),4,cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
      if ((size_t) total > cap)
        return dx_fail_cap(ctx,(size_t) (total),cap);
      a.ticket = d_ticket + 1;
      DX_PROF_BEGIN(ctx); k_qv_code<1><<<grid,kEncThreads,smem1,ctx->stream>>>(a);
      DX_LAUNCHED(ctx,"k_qv_emit");
    }
  if (h_entry_off != NULL)
    { if (max_entries < n)
        return dx_fail(ctx,DX_E_CAP,"entry offset array holds %lld, need %lld",
                       (long long) max_entries,(long long) n);
      DX_CUDA(ctx,cudaMemcpyAsync(h_entry_off,d_off,(size_t) (n+1)*8,cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
    }
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));      // _dev calls return with the work done
  *out_len = (size_t) total;
  if (last_well) *last_well = lastw;
  return DX_OK;
}
