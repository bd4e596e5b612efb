// dx_pack3.cu -- the 2-bit codec for entries whose lines form the usual lattice (every line but the
// last W characters + '\n', W >= 16): the kernels dx_dexta_* / dx_undexta_* use first.
//
// Replaces Number_Read / Number_Arrow + Compress_Read (reference DB.c:393-441, 319-338) behind
// dexta.c:139-205 / dexar.c:138-211, and Uncompress_Read + Lower_/Upper_Read / Letter_Arrow
// (DB.c:342-389) with the line wrapping of undexta.c:263-270 / undexar.c:221-228.
//
// Both kernels are OUTPUT-centric: on the lattice the text position of symbol b is b + b/W, so a
// lane can address the 16 symbols of one 32-bit payload word (pack) or the 32 payload bits behind an
// aligned 16-byte piece of text (unpack) directly.  No warp scans, no bit writer, no staging:
//
//   k_fa_pack3    one warp per entry, one payload word per lane per round.  The 17-byte text window
//                 of the word (16 symbols + at most one newline) comes from three cached 8-byte
//                 loads and funnel shifts, the newline is squeezed out with byte masks, SWAR
//                 compares give the codes, and the word leaves through an aligned 32-bit store (the
//                 payload's byte alignment is absorbed by a shuffle + funnel shift).  Every byte of
//                 the entry's text is checked on the way: symbols must not be '\n' and every lattice
//                 position must hold one, so the symbol count k_fa_measure2 derived from the size of
//                 the entry is PROVEN before the result is used (a mismatch sends the file to the
//                 exact path).
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

// 0x80 flags of the bytes that are '\n', OR-ed over four words (exact per byte)
__device__ __forceinline__ uint32_t newline_flags(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{ return dx_eq_mask(a,'\n') | dx_eq_mask(b,'\n') | dx_eq_mask(c,'\n') | dx_eq_mask(d,'\n'); }

// the low `bits` bits set, bits clamped to 0..32
__device__ __forceinline__ uint32_t low_mask(int bits)
{ return __funnelshift_lc(0xffffffffu,0u,(uint32_t) max(bits,0)); }

__device__ __forceinline__ uint2 ld8_guard(const uint2 *p, const uint8_t *end)
{ return (reinterpret_cast<const uint8_t *>(p + 1) <= end) ? __ldg(p) : make_uint2(0u,0u); }

// ---- pack ----------------------------------------------------------------------------------------------------

struct Pack3Args
{ const uint8_t *text;
  const uint8_t *text_end;      // first address past the readable (16-byte padded) text
  FaEntries ent;
  int32_t lwell_in;
  uint8_t *out;
  int32_t *err;                 // set to 1 when an entry's text is not the lattice it was measured as
  int32_t *leftover;            // set to 1 when an entry was left to k_fa_pack2 (no lattice / W < 16)
  unsigned long long *ticket;
};

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
      if (rlen > 0 && (!(a.ent.flag[e] & 8) || Wi < 16))
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
      const uint32_t W = (uint32_t) Wi;
      const uint32_t clen = ((uint32_t) rlen + 3u) >> 2;
      uint8_t *pay = dst + (a.ent.bytes[e] - clen);
      const uint32_t skew = (uint32_t) (reinterpret_cast<uintptr_t>(pay) & 3u);
      uint32_t *abase = reinterpret_cast<uint32_t *>(pay - skew);
      const uint32_t nw = (clen + skew + 3u) >> 2;                  // aligned words the payload touches
      const uint8_t *seq = a.text + a.ent.seq[e];
      uint32_t b = 16u * (uint32_t) lane;                           // first symbol of my payload word
      uint32_t line = b / W, col = b - line*W;                      // its line and column
      const uint32_t dline = 512u / W, dcol = 512u - dline*W;
      uint32_t carry = 0, bad = 0;
      // the three 8-byte words around the window of the NEXT round are already in flight
      const uint8_t *Pn = seq + b + line;                           // text position of symbol b
      uint2 n0 = make_uint2(0u,0u), n1 = n0, n2 = n0;
      if ((int32_t) b < rlen)
        { const uint2 *A8 = reinterpret_cast<const uint2 *>(reinterpret_cast<uintptr_t>(Pn) & ~(uintptr_t) 7);
          n0 = ld8_guard(A8,a.text_end); n1 = ld8_guard(A8+1,a.text_end); n2 = ld8_guard(A8+2,a.text_end);
        }
#pragma unroll 1
      for (uint32_t j0 = 0; j0 < nw; j0 += 32)
        { const uint32_t j = j0 + (uint32_t) lane;
          const int32_t nv = rlen - (int32_t) b;                    // symbols from b on (16 = a full word)
          const int32_t jn = (int32_t) (W - col);                   // window index of the line's newline
          const uint8_t *P = Pn;
          const uint2 q0 = n0, q1 = n1, q2 = n2;
          b += 512u; line += dline; col += dcol;
          if (col >= W) { col -= W; line++; }
          if ((int32_t) b < rlen)
            { Pn = seq + b + line;
              const uint2 *A8 = reinterpret_cast<const uint2 *>(reinterpret_cast<uintptr_t>(Pn) & ~(uintptr_t) 7);
              n0 = ld8_guard(A8,a.text_end); n1 = ld8_guard(A8+1,a.text_end); n2 = ld8_guard(A8+2,a.text_end);
            }
          uint32_t val = 0;
          if (nv > 0)
            { const uintptr_t PA = reinterpret_cast<uintptr_t>(P);
              const bool up = (PA & 4) != 0;
              const uint32_t x0 = up ? q0.y : q0.x, x1 = up ? q1.x : q0.y, x2 = up ? q1.y : q1.x,
                             x3 = up ? q2.x : q1.y, x4 = up ? q2.y : q2.x;
              const uint32_t sh = (uint32_t) (PA & 3) * 8u;
              // window bytes 0..16: A holds bytes k.., B the same one byte further on
              const uint32_t A0 = __funnelshift_r(x0,x1,sh), A1 = __funnelshift_r(x1,x2,sh),
                             A2 = __funnelshift_r(x2,x3,sh), A3 = __funnelshift_r(x3,x4,sh),
                             A4 = x4 >> sh;
              const uint32_t B0 = __funnelshift_r(A0,A1,8), B1 = __funnelshift_r(A1,A2,8),
                             B2 = __funnelshift_r(A2,A3,8), B3 = __funnelshift_r(A3,A4,8);
              const int32_t jb = (jn > 16) ? 128 : 8*jn;
              const uint32_t M0 = low_mask(jb), M1 = low_mask(jb - 32), M2 = low_mask(jb - 64),
                             M3 = low_mask(jb - 96);
              uint32_t R0 = (A0 & M0) | (B0 & ~M0), R1 = (A1 & M1) | (B1 & ~M1),
                       R2 = (A2 & M2) | (B2 & ~M2), R3 = (A3 & M3) | (B3 & ~M3);
              if (jn <= nv && __ldg(P + jn) != '\n') bad = 1;       // the lattice says: a newline here
              if (nv < 16)                                          // last word of the entry
                { const int vb = 8*nv;
                  R0 &= low_mask(vb); R1 &= low_mask(vb - 32); R2 &= low_mask(vb - 64); R3 &= low_mask(vb - 96);
                }
              if (newline_flags(R0,R1,R2,R3)) bad = 1;              // ... and none among the symbols
              const uint32_t c0 = codes4_top<KIND>(R0), c1 = codes4_top<KIND>(R1),
                             c2 = codes4_top<KIND>(R2), c3 = codes4_top<KIND>(R3);
              val = __byte_perm(__byte_perm(c0,c1,0x0073),__byte_perm(c2,c3,0x0073),0x5410);
              if (KIND == DX_ARROW && nv < 16)                      // the padding is 0, not code('\0') = 3
                { uint32_t keep = 0;
                  for (int m = 0; m < 4; m++)
                    { const int cnt = min(4,max(0,nv - 4*m));
                      keep |= ((0xff00u >> (2*cnt)) & 0xffu) << (8*m);
                    }
                  val &= keep;
                }
            }
          // payload word j-1 | j -> aligned word j
          const uint32_t up1 = __shfl_up_sync(DX_FULL,val,1);
          const uint32_t prev = (lane == 0) ? carry : up1;
          carry = __shfl_sync(DX_FULL,val,31);
          const uint32_t word = __funnelshift_l(prev,val,8u*skew);
          if (j < nw)
            { const int lo = (j == 0) ? (int) skew : 0;
              const int hi = min(4,(int) (clen + skew) - 4*(int) j);
              if (lo == 0 && hi == 4) abase[j] = word;
              else
                { uint8_t *p = reinterpret_cast<uint8_t *>(abase + j);
                  for (int k = lo; k < hi; k++) p[k] = (uint8_t) (word >> (8*k));
                }
            }
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
