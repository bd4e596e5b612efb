// dx_pack3.cu -- the 2-bit codec for entries whose lines form the usual lattice (every line but the
// last W characters + '\n'; W >= 32 to pack, >= 16 to unpack): the kernels dx_dexta_* / dx_undexta_* use first.
//
// Replaces Number_Read / Number_Arrow + Compress_Read (reference DB.c:393-441, 319-338) behind
// dexta.c:139-205 / dexar.c:138-211, and Uncompress_Read + Lower_/Upper_Read / Letter_Arrow
// (DB.c:342-389) with the line wrapping of undexta.c:263-270 / undexar.c:221-228.
//
// On the lattice the text position of symbol b is b + b/W, so both directions can be addressed
// arithmetically.  No warp scans, no bit writer, no staging in shared memory:
//
//   k_fa_pack3    one warp per entry, one ALIGNED 32-byte block of text per lane per round (two
//                 16-byte loads, nothing to realign).  The block's first symbol index and the place
//                 of its (at most one) newline follow from the block's offset; SWAR compares give
//                 the 32 codes as a 64-bit string, the newline's slot is cut out with two masks, the
//                 string is shifted to its bit position in the payload, and the lane stores the
//                 words it completes -- the partial word at its end travels to the next lane by
//                 shuffle.  Every byte of the entry's text is checked on the way: a newline must sit
//                 at every lattice position and nowhere else, so the symbol count k_fa_measure2
//                 derived from the size of the entry is PROVEN before the result is used (a
//                 mismatch sends the file to the exact path).
//   k_unpack3     one warp per entry, one aligned 16-byte store per lane per round: 32 payload bits
//                 from two word loads, the gap for a newline opened with two masks, bit reversal +
//                 three mask-shift steps spread the sixteen 2-bit codes over the nibbles of two
//                 registers, and four PRMTs turn them into characters (the newline is just one more
//                 selector value).
//   k_pk_headers  the header lines of undexta / undexar, one thread per entry.

#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kP3Warps   = 8;
constexpr int kP3Threads = kP3Warps * 32;

// four text bytes -> their four 2-bit codes in bits 31..24, first symbol on top (DB.c:333-334)
template <int KIND>
__device__ __forceinline__ uint32_t codes4_top(uint32_t w)
{ uint32_t b0, b1;                             // 0x80 per byte: bit 0 / bit 1 of the code
  if (KIND == DX_FASTA)                        // DB.c:394-411: acgt / ACGT -> 0123, everything else 0
    { const uint32_t x = w | 0x20202020u;
      const uint32_t mc = dx_eq_mask(x,'c'), mg = dx_eq_mask(x,'g'), mt = dx_eq_mask(x,'t');
      b0 = mc | mt; b1 = mg | mt;
    }
  else                                         // DB.c:419-436: '1','2','3' -> 0,1,2 ; 'G' -> 2 ; else 3
    { const uint32_t m1 = dx_eq_mask(w,'1'), m2 = dx_eq_mask(w,'2');
      const uint32_t m3 = dx_eq_mask(w,'3') | dx_eq_mask(w,'G');
      b0 = ~(m1 | m3) & 0x80808080u; b1 = ~(m1 | m2) & 0x80808080u;
    }
  const uint32_t v = (b0 >> 7) | (b1 >> 6);
  return v * 0x40100401u;                      // byte k's code lands at bits 31-2k..30-2k
}

// the low `bits` bits set, bits clamped to 0..32
__device__ __forceinline__ uint32_t low_mask(int bits)
{ return __funnelshift_lc(0xffffffffu,0u,(uint32_t) max(bits,0)); }

// ---- pack ----------------------------------------------------------------------------------------------------

struct Pack3Args
{ const uint8_t *text;
  const uint8_t *text_end;      // first address past the readable (16-byte padded) text
  FaEntries ent;
  int32_t lwell_in;
  uint8_t *out;
  int32_t *err;                 // set to 1 when an entry's text is not the lattice it was measured as
  int32_t *leftover;            // set to 1 when an entry was left to k_fa_pack2 (no lattice / W < 32)
  unsigned long long *ticket;
};

// the top `k` bits set, k clamped to 0..32
__device__ __forceinline__ uint32_t top_mask(int k)
{ return __funnelshift_rc(0u,0xffffffffu,(uint32_t) max(k,0)); }

// aligned payload word g of an entry <- x (its first byte in the top bits); the first and the last
// word may be shared with the neighbouring header fields / the next entry, so only their payload
// bytes are written
__device__ __forceinline__ void store_word(uint32_t *abase, uint32_t g, uint32_t x, uint32_t skew, uint32_t pend)
{ const uint32_t v = __byte_perm(x,0u,0x0123);
  const int lo = (g == 0) ? (int) skew : 0;
  const int hi = min(4,(int) pend - 4*(int) g);                  // pend = skew + payload bytes
  if (lo == 0 && hi == 4) abase[g] = v;
  else
    { uint8_t *p = reinterpret_cast<uint8_t *>(abase + g);
      for (int k = lo; k < hi; k++) p[k] = (uint8_t) (v >> (8*k));
    }
}

// One warp per entry, one ALIGNED 32-byte block of text per lane per round (two 16-byte loads, no
// realignment).  On the lattice the block's first symbol index and the place of its (at most one)
// newline follow from the block's offset; the 32 codes form a 64-bit string from which the newline's
// slot is cut, and the string is shifted to its bit position in the payload.  A lane stores the words
// it completes; the partial word at its end goes to the next lane by shuffle (to lane 0 of the next
// round in a register).
template <int KIND>
__global__ void __launch_bounds__(kP3Threads)
k_fa_pack3(Pack3Args a)
{ const int lane = threadIdx.x & 31;
  const uint32_t fields = (KIND == DX_FASTA) ? 12u : 16u;
  unsigned long long next = 0;
  if (lane == 0) next = atomicAdd(a.ticket,1ull);
  while (true)
    { const int64_t e = (int64_t) __shfl_sync(DX_FULL,next,0);
      if (e >= a.ent.n) break;
      if (lane == 0) next = atomicAdd(a.ticket,1ull);
      const int32_t rlen = a.ent.rlen[e];
      const int32_t Wi   = a.ent.width[e];
      if (rlen > 0 && (!(a.ent.flag[e] & 8) || Wi < 32))
        { if (lane == 0) atomicExch(a.leftover,1);
          continue;
        }
      uint8_t *dst = a.out + a.ent.off[e];
      if (lane == 0)
        { // entry header: well-delta bytes, beg, end, qv | 4 x uint16 SNR (dexta.c:187-198)
          int32_t lwell = (e == 0) ? a.lwell_in : a.ent.well[e-1];
          const int32_t well = a.ent.well[e];
          uint8_t *h = dst;
          while (well - lwell >= 255) { *h++ = 0xff; lwell += 255; }
          *h++ = (uint8_t) (well - lwell);
          const uint32_t f[4] = { (uint32_t) a.ent.beg[e], (uint32_t) a.ent.end[e],
                                  (uint32_t) a.ent.aux[2*e], (uint32_t) a.ent.aux[2*e+1] };
          for (uint32_t k = 0; k < fields; k++)
            *h++ = (uint8_t) (f[k >> 2] >> (8*(k & 3)));
        }
      if (rlen <= 0) continue;
      const uint32_t W = (uint32_t) Wi, Wp1 = W + 1u;
      const uint32_t clen = ((uint32_t) rlen + 3u) >> 2;
      uint8_t *pay = dst + (a.ent.bytes[e] - clen);
      const uint32_t skew = (uint32_t) (reinterpret_cast<uintptr_t>(pay) & 3u);
      uint32_t *abase = reinterpret_cast<uint32_t *>(pay - skew);
      const uint32_t pend = clen + skew;
      const uint8_t *seq = a.text + a.ent.seq[e];
      const int32_t rend = (int32_t) a.ent.region[e] - 1;           // the final newline (k_fa_measure2 saw it)
      const int32_t d0 = (int32_t) (reinterpret_cast<uintptr_t>(seq) & 31);
      const uint8_t *blk = seq - d0 + 32*lane;                      // my block of the round
      int32_t t0 = 32*lane - d0;                                    // text offset of its byte 0
      int32_t line; uint32_t col;                                   // t0 = line*(W+1) + col, 0 <= col <= W
      if (t0 >= 0) { line = (int32_t) ((uint32_t) t0 / Wp1); col = (uint32_t) t0 - (uint32_t) line*Wp1; }
      else         { line = -1; col = Wp1 - (uint32_t) (-t0); }
      const uint32_t dline = 1024u / Wp1, dcol = 1024u - dline*Wp1;
      uint32_t carry = 0, bad = 0;
#pragma unroll 1
      for (int32_t r0 = -d0; r0 < rend; r0 += 1024)                 // r0 = t0 of lane 0
        { uint32_t X0 = 0, X1 = 0, X2 = 0, f = 0, nstore = 0, cout = 0;
          if (t0 < rend)
            { uint4 q0 = make_uint4(0,0,0,0), q1 = q0;
              if (blk >= a.text && blk < a.text_end) q0 = dx_ldg16(blk);
              if (blk + 16 >= a.text && blk + 16 < a.text_end) q1 = dx_ldg16(blk + 16);
              uint32_t w[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
              const int32_t vlo = max(0,-t0), vhi = min(32,rend - t0);
              const uint32_t ln = (t0 < 0) ? 0u : (uint32_t) line, cl = (t0 < 0) ? 0u : col;
              if (vlo > 0 || vhi < 32)                              // first / last block: blank what is not mine
                {
#pragma unroll
                  for (int k = 0; k < 8; k++)
                    w[k] &= low_mask(8*(vhi - 4*k)) & ~low_mask(8*(vlo - 4*k));
                }
              const uint32_t b0 = ln*W + cl;                        // symbols before my first byte
              const int32_t jn = vlo + (int32_t) (W - cl);          // block index of the line's newline
              const bool hasnl = (jn < vhi);
              // every newline must be the lattice's, and the lattice's must be there
              const uint32_t ebit = hasnl ? (0x80u << (8*(jn & 3))) : 0u;
              const int32_t  ek   = jn >> 2;
              uint32_t c[8];
#pragma unroll
              for (int k = 0; k < 8; k++)
                { bad |= dx_eq_mask(w[k],'\n') ^ ((k == ek) ? ebit : 0u);
                  c[k] = codes4_top<KIND>(w[k]);
                }
              // symbol i of the block at bits 63-2i, 62-2i of hi:lo
              uint32_t hi = __byte_perm(__byte_perm(c[2],c[3],0x0037),__byte_perm(c[0],c[1],0x3700),0x7610);
              uint32_t lo = __byte_perm(__byte_perm(c[6],c[7],0x0037),__byte_perm(c[4],c[5],0x3700),0x7610);
              if (hasnl)                                            // cut the newline's slot out
                { const uint32_t mh = top_mask(2*jn), ml = top_mask(2*jn - 32);
                  const uint32_t h2 = __funnelshift_l(lo,hi,2), l2 = lo << 2;
                  hi = (hi & mh) | (h2 & ~mh);
                  lo = (lo & ml) | (l2 & ~ml);
                }
              const int32_t nsym = (vhi - vlo) - (hasnl ? 1 : 0);
              if (vlo > 0 || vhi < 32)
                { if (vlo >= 16) { hi = lo << (2*vlo - 32); lo = 0; }
                  else if (vlo > 0) { hi = __funnelshift_l(lo,hi,2*vlo); lo <<= 2*vlo; }
                  hi &= top_mask(2*nsym); lo &= top_mask(2*nsym - 32);
                }
              const uint32_t nbits = 2u * (uint32_t) nsym;
              const uint32_t s = 2u*b0 + 8u*skew;                   // my bit position in the payload words
              const uint32_t o = s & 31u;
              f = s >> 5;
              X0 = hi >> o; X1 = __funnelshift_r(lo,hi,o); X2 = __funnelshift_r(0u,lo,o);
              const uint32_t endb = s + nbits;                      // one past my last bit
              nstore = (endb >> 5) - f;                             // words I complete (0..2)
              if (endb & 31u)                                       // ... and a partial one at the end
                { const uint32_t part = (nstore == 0) ? X0 : (nstore == 1) ? X1 : X2;
                  if (t0 + 32 >= rend) nstore++;                    // the entry ends here: the rest is padding
                  else
                    { // a block in the middle holds >= 62 bits, so only the entry's first block (lane 0,
                      // whose predecessor is the carry register) can end inside the word it starts in
                      cout = part;
                      if (nstore == 0 && lane == 0) cout |= carry;
                    }
                }
            }
          // the partial word my predecessor ended with is the start of my first word
          uint32_t cin = __shfl_up_sync(DX_FULL,cout,1);
          if (lane == 0) cin = carry;
          carry = __shfl_sync(DX_FULL,cout,31);
          if (t0 < rend)
            { X0 |= cin;
              if (nstore > 0) store_word(abase,f,X0,skew,pend);
              if (nstore > 1) store_word(abase,f+1,X1,skew,pend);
              if (nstore > 2) store_word(abase,f+2,X2,skew,pend);
            }
          blk += 1024; t0 += 1024; line += (int32_t) dline; col += dcol;
          if (col >= Wp1) { col -= Wp1; line++; }
        }
      if (__any_sync(DX_FULL,bad != 0) && lane == 0) atomicExch(a.err,1);
    }
}

// ---- unpack --------------------------------------------------------------------------------------------------

__device__ int fmt_int3(uint8_t *p, int32_t v)
{ char tmp[12];
  int  k = 0, len = 0;
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  if (v < 0) p[len++] = '-';
  do { tmp[k++] = (char) ('0' + u % 10); u /= 10; } while (u);
  while (k) p[len++] = (uint8_t) tmp[--k];
  return len;
}

// "%s/%d/%d_%d RQ=0.%d\n" (undexta.c:242) or "%s/%d/%d_%d SN=%.2f,%.2f,%.2f,%.2f\n" (undexar.c:202)
__global__ void __launch_bounds__(128)
k_pk_headers(int kind, const PkDecEntry *ent, int64_t count, const char *prefix, int plen, uint8_t *out)
{ const int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  const PkDecEntry en = ent[e];
  uint8_t *h = out + en.out_off;
  int hl = 0;
  for (int k = 0; k < plen; k++) h[hl++] = (uint8_t) prefix[k];
  h[hl++] = '/'; hl += fmt_int3(h+hl,en.well);
  h[hl++] = '/'; hl += fmt_int3(h+hl,en.beg);
  h[hl++] = '_'; hl += fmt_int3(h+hl,en.end);
  if (kind == DX_FASTA)
    { const char *rq = " RQ=0.";
      for (int k = 0; k < 6; k++) h[hl++] = (uint8_t) rq[k];
      hl += fmt_int3(h+hl,en.aux[0]);
    }
  else
    { const char *sn = " SN=";
      for (int k = 0; k < 4; k++) h[hl++] = (uint8_t) sn[k];
      for (int k = 0; k < 4; k++)
        { const uint32_t c = ((uint32_t) en.aux[k >> 1] >> (16*(k & 1))) & 0xffffu;
          hl += fmt_int3(h+hl,(int32_t) (c / 100u));
          h[hl++] = '.';
          h[hl++] = (uint8_t) ('0' + (c % 100u) / 10u);
          h[hl++] = (uint8_t) ('0' + c % 10u);
          if (k < 3) h[hl++] = ',';
        }
    }
  h[hl++] = '\n';
}

struct Unpack3Args
{ int kind, upper, width;
  const uint8_t *in; const uint8_t *in_end4;         // first 4-byte aligned address past the readable image
  const PkDecEntry *ent; int64_t count;
  uint8_t *out;
  unsigned long long *ticket;
};

// 16 codes, one per 2 bits from bit 0 up -> one per nibble of (lo, hi)
__device__ __forceinline__ uint32_t spread16(uint32_t h)
{ uint32_t y = (h | (h << 8)) & 0x00ff00ffu;
  y = (y | (y << 4)) & 0x0f0f0f0fu;
  return (y | (y << 2)) & 0x33333333u;
}

// the payload bits behind the 16 characters at text offsets t .. t+15 of an entry, t = line*(W+1) + col
struct Bits16 { uint32_t w0, w1, sh, j; };

__device__ __forceinline__ Bits16 load16(const Unpack3Args &a, const uint8_t *pay, uint32_t line, uint32_t col,
                                         uint32_t W)
{ Bits16 r;
  const uint32_t b = line*W + col;                               // first symbol at or after t
  const uintptr_t A = reinterpret_cast<uintptr_t>(pay + (b >> 2));
  const uint32_t *a4 = reinterpret_cast<const uint32_t *>(A & ~(uintptr_t) 3);
  r.w0 = (reinterpret_cast<const uint8_t *>(a4) < a.in_end4) ? __ldg(a4) : 0u;
  r.w1 = (reinterpret_cast<const uint8_t *>(a4 + 1) < a.in_end4) ? __ldg(a4 + 1) : 0u;
  r.sh = (uint32_t) (A & 3)*8u + (b & 3u)*2u;
  r.j  = W - col;                                                // where the line's newline falls
  return r;
}

// ... as characters (the caller patches the final newline, the only one that may be off the lattice)
__device__ __forceinline__ uint4 text16(const Bits16 &q, uint32_t alpha)
{ uint32_t x = __funnelshift_l(__byte_perm(q.w1,0,0x0123),__byte_perm(q.w0,0,0x0123),q.sh);   // symbol k at bits 31-2k..
  const uint32_t j = q.j;
  if (j < 16u)                                                   // open a 2-bit gap for the newline
    { const uint32_t m = ~(0xffffffffu >> (2u*j));
      x = (x & m) | ((x & ~m) >> 2);
    }
  const uint32_t r = __brev(x);                                  // symbol k at bits 2k+1..2k, its two bits swapped
  uint32_t ylo = spread16(r & 0xffffu), yhi = spread16(r >> 16);
  if (j < 16u)
    { const uint32_t nl = 4u << (4u*(j & 7u));                   // selector 4: byte 0 of the second operand
      if (j < 8u) ylo |= nl; else yhi |= nl;
    }
  const uint32_t nl4 = 0x0a0a0a0au;
  return make_uint4(__byte_perm(alpha,nl4,ylo),__byte_perm(alpha,nl4,ylo >> 16),
                    __byte_perm(alpha,nl4,yhi),__byte_perm(alpha,nl4,yhi >> 16));
}

__global__ void __launch_bounds__(kP3Threads)
k_unpack3(Unpack3Args a)
{ const int lane = threadIdx.x & 31;
  // indexed by the bit-swapped code (see text16): 0 -> a, 1 (= code 2) -> g, 2 (= code 1) -> c, 3 -> t
  const uint32_t alpha = (a.kind == DX_ARROW) ? 0x34323331u          // "1324"
                        : a.upper ? 0x54434741u : 0x74636761u;       // "AGCT" / "agct"
  const uint32_t W = (uint32_t) a.width, Wp1 = W + 1u;
  const uint32_t dline = 512u / Wp1, dcol = 512u - dline*Wp1;
  unsigned long long next = 0;
  if (lane == 0) next = atomicAdd(a.ticket,1ull);
  while (true)
    { const int64_t e = (int64_t) __shfl_sync(DX_FULL,next,0);
      if (e >= a.count) break;
      if (lane == 0) next = atomicAdd(a.ticket,1ull);
      const PkDecEntry en = a.ent[e];
      const int64_t rl64 = (int64_t) en.end - en.beg;
      if (rl64 <= 0) continue;
      const uint32_t rlen = (uint32_t) rl64;
      // text of the entry: rlen symbols, a '\n' after every W of them and after the last
      const uint32_t tlen = rlen + (rlen + W - 1u) / W;
      const uint8_t *pay = a.in + en.bin_off;
      uint8_t *dst = a.out + en.text_off;
      const int32_t skew = (int32_t) (reinterpret_cast<uintptr_t>(dst) & 15);
      uint8_t *base = dst - skew;                                    // 16-byte aligned
      const uint32_t nchunk = ((uint32_t) skew + tlen + 15u) >> 4;
      int32_t t0 = 16*lane - skew;                                   // text offset of my chunk's byte 0
      int32_t line; uint32_t col;                                    // t0 = line*(W+1) + col, 0 <= col <= W
      if (t0 >= 0) { line = (int32_t) ((uint32_t) t0 / Wp1); col = (uint32_t) t0 - (uint32_t) line*Wp1; }
      else         { line = -1; col = Wp1 - (uint32_t) (-t0); }
      // the chunk that starts before the entry's text (lane 0, skew > 0) is made from offset 0 on
      Bits16 nxt;
      if ((uint32_t) lane < nchunk) nxt = (t0 < 0) ? load16(a,pay,0u,0u,W) : load16(a,pay,(uint32_t) line,col,W);
#pragma unroll 1
      for (uint32_t c = (uint32_t) lane; c < nchunk; c += 32)
        { const Bits16 cur = nxt;
          const int32_t tc = t0;
          t0 += 512; line += (int32_t) dline; col += dcol;
          if (col >= Wp1) { col -= Wp1; line++; }
          if (c + 32 < nchunk && (uint32_t) t0 < tlen) nxt = load16(a,pay,(uint32_t) line,col,W);   // in flight
          const uint4 o = text16(cur,alpha);
          if (tc >= 0 && (uint32_t) tc + 17u <= tlen)
            dx_stg16(base + (size_t) c*16,o);                        // 16 lattice positions before the final newline
          else
            { // first / last chunk: the final newline patched in, stored byte by byte where the
              // characters belong to this entry
              const uint32_t tv = (tc < 0) ? 0u : (uint32_t) tc;
              const uint32_t w[4] = { o.x, o.y, o.z, o.w };
              const int lo = (tc < 0) ? -tc : 0, hi = (int) min(16u,(uint32_t) ((int32_t) tlen - tc));
              for (int k = lo; k < hi; k++)
                { const uint32_t t = (uint32_t) (tc + k), i = t - tv;
                  const uint32_t ww = (i & 8u) ? ((i & 4u) ? w[3] : w[2]) : ((i & 4u) ? w[1] : w[0]);
                  uint32_t ch = (ww >> (8u*(i & 3u))) & 0xffu;
                  if (t == tlen - 1u) ch = '\n';
                  base[(size_t) c*16 + k] = (uint8_t) ch;
                }
            }
        }
    }
}

}  // namespace

int dxk_fa_pack3(dx_ctx *ctx, int kind, const uint8_t *d_text, size_t n, FaEntries ent, int32_t lwell_in,
                 uint8_t *d_out, int32_t *d_err, int32_t *d_leftover, unsigned long long *d_ticket)
{ if (ent.n == 0) return DX_OK;
  Pack3Args a;
  a.text = d_text; a.text_end = d_text + ((n + 15) & ~(size_t) 15);
  a.ent = ent; a.lwell_in = lwell_in; a.out = d_out; a.err = d_err; a.leftover = d_leftover; a.ticket = d_ticket;
  int64_t grid = (ent.n + kP3Warps - 1) / kP3Warps;
  if (grid > (int64_t) ctx->sm_count * 8) grid = (int64_t) ctx->sm_count * 8;
  DX_PROF_BEGIN(ctx);
  if (kind == DX_FASTA) k_fa_pack3<DX_FASTA><<<(unsigned) grid,kP3Threads,0,ctx->stream>>>(a);
  else                  k_fa_pack3<DX_ARROW><<<(unsigned) grid,kP3Threads,0,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,"k_fa_pack3");
  return DX_OK;
}

int dxk_unpack3(dx_ctx *ctx, int kind, int upper, int width, const uint8_t *d_in, size_t n, const PkDecEntry *d_ent,
                int64_t count, const char *d_prefix, int plen, uint8_t *d_out, unsigned long long *d_ticket)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx);
  k_pk_headers<<<(unsigned) ((count + 127)/128),128,0,ctx->stream>>>(kind,d_ent,count,d_prefix,plen,d_out);
  DX_LAUNCHED(ctx,"k_pk_headers");
  Unpack3Args a;
  a.kind = kind; a.upper = upper; a.width = width; a.in = d_in;
  a.in_end4 = d_in + ((n + 15) & ~(size_t) 15);
  a.ent = d_ent; a.count = count; a.out = d_out; a.ticket = d_ticket;
  int64_t grid = (count + kP3Warps - 1) / kP3Warps;
  if (grid > (int64_t) ctx->sm_count * 8) grid = (int64_t) ctx->sm_count * 8;
  DX_PROF_BEGIN(ctx); k_unpack3<<<(unsigned) grid,kP3Threads,0,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,"k_unpack3");
  return DX_OK;
}
