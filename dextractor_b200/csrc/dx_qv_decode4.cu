// dx_qv_decode4.cu -- parallel .dexqv entry decoder (the one the library uses).
//
// Replaces Decode / Decode_Run (reference QV.c:510-691) + Packed_Length / Unpack_Tag
// (QV.c:823-847) + the per-entry text output of undexqv.c:182-207.
//
// One CTA per entry, every stream decoded by all threads at once.  A Huffman stream cannot be cut
// at known code boundaries, so it is cut into fixed 256-bit subsequences that are decoded
// speculatively (prefix codes resynchronise after a few symbols):
//
//   0. the window's words are staged once in shared memory as overlapping 64-bit pairs (word j in
//      the high half, word j+1 in the low half), already shifted to the stream's byte alignment:
//      the 32 bits at ANY bit position are one 64-bit shared load and one funnel shift.  Next to
//      them sit the stream's 12-bit decode tables (plain streams: up to TWO symbols per lookup).
//      Only as many subsequences as the code lengths predict for the symbols still to come are
//      active; whatever is left of the stream simply becomes the next window;
//   1. thread i decodes from bit 256*i (a guess) to the first code boundary at or past bit
//      256*(i+1), its EXIT, counting symbols;
//   2. rounds: a thread whose start differs from its predecessor's exit walks two fingers, one
//      from its old start and one from the new one, always advancing the one behind by a single
//      symbol, until they meet: from there on both paths are the same, so exit and counts follow
//      by arithmetic (no re-decode of the rest).  Thread 0 starts at the true position, so at the
//      fix point every start is a true code boundary (induction over i);
//   3. an exclusive scan of the symbol counts places every subsequence in the output line and
//      finds the subsequence in which the rlen-th symbol -- hence the stream -- ends; the stream's
//      length in the file follows from the position of its last item ((p_last+47)>>5 words, the
//      reference's refill rule, QV.c:537-551);
//   4. every thread decodes its subsequence once more, now producing text: run-length streams
//      scatter their non-run symbols into a line pre-filled with the run character, plain streams
//      go through a shared-memory stage that is flushed with aligned 32-bit stores.
// Speculation only costs time: nothing is written before the fix point is reached.

#include <stdio.h>
#include <stdlib.h>
#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kT        = 256;                  // threads per CTA = subsequences per window
constexpr int kS        = 256;                  // bits per subsequence
constexpr int kSW       = kS / 32;              // words per subsequence
constexpr int kWinWords = kT * kSW;
constexpr int kPadWords = 8;                    // look-ahead of the last subsequence
constexpr int kOutStage = 16384;                // bytes of the output stage of plain streams
constexpr int kTail     = 24;                   // longest single step (16-bit code + 8-bit literal)

struct Dec4Args
{ const uint8_t *in;
  int64_t        n;
  const QvDecTables4 *tab;
  int32_t        delchar, subchar, upper, write;
  int64_t        count;
  const int64_t *start;        // first stream byte of each entry (after beg/end/qv)
  const int32_t *rlen;
  const QvDecEntry *ent;       // write mode: output placement
  const char    *prefix; int32_t plen;
  uint8_t       *out;
  int64_t       *soff;         // [count][6] or NULL
  int32_t       *status;       // [count] (walk) or [1] (decode)
  unsigned long long *ticket;
  unsigned long long *dbg;     // optional counters [table][0 rounds, 1 windows, 2 streams, 3 restarts]
};

struct Shared4
{ uint64_t bits[kWinWords + kPadWords];     // entry j = stream word j << 32 | stream word j+1
  union
  { uint32_t multi[4096];                   // plain stream: up to two symbols per entry
    struct { uint16_t run[4096], sym[4096]; } rs;
  } tb;
  union
  { uint32_t stage[kOutStage/4];
    uint32_t tagstage[kT/32][132];
  } u;
  uint32_t exitst[2][kT];                   // exit state of every subsequence (double buffered)
  uint32_t wsum[kT/32];
  uint32_t end_words, bad, kept_sum;
  int64_t  entry;
};

extern __shared__ __align__(16) uint8_t dx_dec4_smem[];
#define DX_SM (*reinterpret_cast<Shared4 *>(dx_dec4_smem))

// the 32 stream bits that start at window-relative bit `pos`
__device__ __forceinline__ uint32_t win32(const uint64_t *D, uint32_t pos)
{ const uint64_t v = D[pos >> 5];
  return __funnelshift_l((uint32_t) v,(uint32_t) (v >> 32),pos);
}

// codes longer than 12 bits: sym | len << 8, len 0 = no code maps here
__device__ __forceinline__ uint32_t lookup_long(const QvDecTables2 *t, int k, uint32_t w16)
{ uint32_t e = __ldg(&t->prim[k][w16 >> 5]);
  if (e & 0x8000u)
    e = __ldg(&t->sub[k][(e & 0x7fffu)*32u + (w16 & 31u)]);
  return e;
}

// ... as an entry of the plain-stream table (see QvDecTables4)
__device__ __noinline__ uint32_t long_entry(const QvDecTables2 *t, int k, uint32_t w, uint32_t *bad)
{ const uint32_t f = lookup_long(t,k,w >> 16);
  uint32_t len = (f >> 8) & 31u;
  const uint32_t c = f & 0xffu;
  if (len == 0) { len = 1; *bad = 1; }
  if (t->type[k] == 2 && c == 255u)
    return (len + 8u) | (1u << 5) | 0x80u | (len << 8) | (255u << 16);
  return len | (1u << 5) | (len << 8) | (c << 16);
}

// ... as an entry of a single-symbol table: sym | len << 8, len >= 1
__device__ __noinline__ uint32_t long_single(const QvDecTables2 *t, int k, uint32_t w, uint32_t *bad)
{ uint32_t f = lookup_long(t,k,w >> 16) & 0x1fffu;
  if ((f >> 8) == 0u) { f |= 0x100u; *bad = 1; }
  return f;
}

// block-wide exclusive scan of v (one value per thread); returns exclusive prefix, total in *tot
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *tot)
{ Shared4 &sm = DX_SM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = dx_warp_incl_sum(v,lane);
  if (lane == 31) sm.wsum[warp] = inc;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kT/32; w++)
    { const uint32_t x = sm.wsum[w];
      if (w < warp) before += x;
      all += x;
    }
  __syncthreads();
  *tot = all;
  return before + inc - v;
}

// the CTA fills dst[0..n) with byte c (any alignment)
__device__ __forceinline__ void fill_line(uint8_t *dst, uint32_t c, uint32_t n)
{ const uint32_t t = threadIdx.x;
  uint32_t head = (16u - (uint32_t) (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > n) head = n;
  if (t < head) dst[t] = (uint8_t) c;
  const uint32_t nvec = (n - head) >> 4;
  const uint32_t q = c * 0x01010101u;
  const uint4 v = make_uint4(q,q,q,q);
  uint8_t *body = dst + head;
  for (uint32_t i = t; i < nvec; i += kT) dx_stg16(body + (size_t) i*16,v);
  const uint32_t done = head + nvec*16u;
  if (t < n - done) dst[done + t] = (uint8_t) c;
}

// stage `nwords` stream words: word j = the LE uint32 at byte p + 4j (zeros past the image)
__device__ __forceinline__ void stage_window(const Dec4Args &a, const uint8_t *p, int nwords)
{ Shared4 &sm = DX_SM;
  const uintptr_t A = reinterpret_cast<uintptr_t>(p);
  const uint32_t *al = reinterpret_cast<const uint32_t *>(A & ~(uintptr_t) 3);
  const uint32_t sh = (uint32_t) (A & 3) * 8;
  const int64_t limit = ((int64_t) (reinterpret_cast<uintptr_t>(a.in + a.n) + 3) -
                         (int64_t) (A & ~(uintptr_t) 3)) >> 2;         // aligned words readable
  for (int j = threadIdx.x; j < nwords; j += kT)
    { const uint32_t x0 = (j     < limit) ? __ldg(al + j)     : 0u;
      const uint32_t x1 = (j + 1 < limit) ? __ldg(al + j + 1) : 0u;
      const uint32_t x2 = (j + 2 < limit) ? __ldg(al + j + 2) : 0u;
      const uint32_t w0 = __funnelshift_r(x0,x1,sh), w1 = __funnelshift_r(x1,x2,sh);
      sm.bits[j] = ((uint64_t) w0 << 32) | w1;
    }
}

// multi entry: bits 0-4 total length (escape: code + 8 literal bits), 5-6 symbols (1|2),
// bit 7 escape, 8-12 length of the first code, 16-23 first symbol, 24-31 second symbol; 0 = long
#define DX_E_LEN(e)   ((e) & 31u)
#define DX_E_N(e)     (((e) >> 5) & 3u)
#define DX_E_LEN0(e)  (((e) >> 8) & 31u)
#define DX_E_LEN1(e)  (((e) & 0x40u) ? DX_E_LEN0(e) : DX_E_LEN(e))      /* one symbol only */

// Decode one stream of `rlen` symbols that starts at byte `so`.  Returns the number of bytes the
// stream occupies; *kept_out = symbol items != rc.  When `dst` is not NULL the line is written.
template <bool RUN>
__device__ __noinline__ uint32_t decode_stream(const Dec4Args &a, int64_t so, int32_t rlen,
                                               int symtab, int runtab, int32_t rci, uint8_t *dst,
                                               uint32_t *kept_out)
{ Shared4 &sm = DX_SM;
  const int t = threadIdx.x;
  *kept_out = 0;
  if (rlen <= 0) return 0;
  const QvDecTables2 *t2 = &a.tab->t2;
  const bool esc = (t2->type[symtab] == 2);
  const uint32_t rc = (uint32_t) rci;                 // 0xffffffff for plain streams
  const uint32_t base = (uint32_t) t * kS, lim = base + kS;
  const float abits = a.tab->abits[symtab] * 1.2f;
  const uint64_t *D = sm.bits;
  const uint32_t *mt = sm.tb.multi;
  const uint16_t *rt = sm.tb.rs.run, *st = sm.tb.rs.sym;
  uint32_t done = 0;                                  // symbols placed by earlier windows
  uint32_t carry = 0;                                 // start state of thread 0 (window relative)
  uint32_t wword = 0;                                 // first stream word of the window
  uint32_t words = 0;
  if (t == 0) sm.kept_sum = 0;
  if (RUN && dst != NULL) fill_line(dst,rc,(uint32_t) rlen);

  // the stream's tables
  if (RUN)
    { const uint32_t *gr = reinterpret_cast<const uint32_t *>(a.tab->single[runtab]);
      const uint32_t *gs = reinterpret_cast<const uint32_t *>(a.tab->single[symtab]);
      uint32_t *sr = reinterpret_cast<uint32_t *>(sm.tb.rs.run);
      uint32_t *ss = reinterpret_cast<uint32_t *>(sm.tb.rs.sym);
      for (int j = t; j < 2048; j += kT) { sr[j] = __ldg(gr + j); ss[j] = __ldg(gs + j); }
    }
  else
    { const uint32_t *gm = a.tab->multi[symtab];
      for (int j = t; j < 4096; j += kT) sm.tb.multi[j] = __ldg(gm + j);
    }

  while (true)
    { const uint32_t remaining = (uint32_t) rlen - done;
      uint32_t nact = (uint32_t) ((float) remaining * abits * (1.0f/kS)) + 3u;
      if (nact > (uint32_t) kT) nact = kT;
      const bool active = ((uint32_t) t < nact);
      stage_window(a,a.in + so + (int64_t) wword*4,(int) nact*kSW + kPadWords);
      __syncthreads();

      // ---- round 0: from the guessed start to the exit --------------------------------------------
      // state = bit position << 1 | parity (1: a run item was read, its symbol item comes next)
      uint32_t mystart = (t == 0) ? carry : (base << 1);
      uint32_t myexit = 0, n = 0, nk = 0, bd = 0;
      if (active)
        { uint32_t pos = mystart >> 1, cnt = 0, kept = 0;
          if (RUN)
            { uint32_t par = mystart & 1u;
              while (pos < lim)
                { if (par == 0)
                    { uint32_t w = win32(D,pos);
                      uint32_t e = rt[w >> 20];
                      if (e == 0u) e = long_single(t2,runtab,w,&bd);
                      uint32_t r = e & 0xffu;
                      pos += e >> 8;
                      if (r == 255u) { r = win32(D,pos) >> 16; pos += 16; }
                      cnt += r;
                      par = 1;
                      if (pos >= lim) break;
                    }
                  uint32_t w = win32(D,pos);
                  uint32_t e = st[w >> 20];
                  if (e == 0u) e = long_single(t2,symtab,w,&bd);
                  uint32_t c = e & 0xffu;
                  pos += e >> 8;
                  if (esc && c == 255u) { c = win32(D,pos) >> 24; pos += 8; }
                  cnt += 1;
                  kept += (c != rc);
                  par = 0;
                }
              myexit = (pos << 1) | par;
            }
          else
            { const uint32_t limf = lim - kTail;      // below it no step can reach the limit
              while (pos < limf)
                { const uint32_t w = win32(D,pos);
                  uint32_t e = mt[w >> 20];
                  if (e == 0u) e = long_entry(t2,symtab,w,&bd);
                  pos += DX_E_LEN(e);
                  cnt += DX_E_N(e);
                }
              while (pos < lim)                       // one symbol at a time: the exit is the FIRST
                { const uint32_t w = win32(D,pos);    // code boundary at or past the limit
                  uint32_t e = mt[w >> 20];
                  if (e == 0u) e = long_entry(t2,symtab,w,&bd);
                  pos += DX_E_LEN1(e);
                  cnt += 1;
                }
              myexit = pos << 1;
            }
          n = cnt; nk = kept;
        }
      int cur = 0;
      sm.exitst[0][t] = myexit;
      __syncthreads();

      // ---- rounds: adopt the predecessor's exit; two fingers until the old path is met ----------
      uint32_t rounds = 0, restarts = 0;
      while (true)
        { int changed = 0;
          const uint32_t want = (active && t > 0) ? sm.exitst[cur][t-1] : mystart;
          if (want != mystart)
            { uint32_t pa = mystart >> 1, pb = want >> 1, ca = 0, cb = 0, ka = 0, kb = 0;
              restarts++;
              if (RUN)
                { uint32_t qa = mystart & 1u, qb = want & 1u;
                  while (!(pa == pb && qa == qb) && min(pa,pb) < lim)
                    { const bool fa = (pa <= pb);
                      uint32_t pos = fa ? pa : pb, par = fa ? qa : qb, dc, dk = 0;
                      if (par == 0)
                        { uint32_t w = win32(D,pos);
                          uint32_t e = rt[w >> 20];
                          if (e == 0u) e = long_single(t2,runtab,w,&bd);
                          dc = e & 0xffu;
                          pos += e >> 8;
                          if (dc == 255u) { dc = win32(D,pos) >> 16; pos += 16; }
                          par = 1;
                        }
                      else
                        { uint32_t w = win32(D,pos);
                          uint32_t e = st[w >> 20];
                          if (e == 0u) e = long_single(t2,symtab,w,&bd);
                          uint32_t c = e & 0xffu;
                          pos += e >> 8;
                          if (esc && c == 255u) { c = win32(D,pos) >> 24; pos += 8; }
                          dc = 1; dk = (c != rc);
                          par = 0;
                        }
                      if (fa) { pa = pos; qa = par; ca += dc; ka += dk; }
                      else    { pb = pos; qb = par; cb += dc; kb += dk; }
                    }
                  if (pa == pb && qa == qb) { n += cb - ca; nk += kb - ka; }
                  else { n = cb; nk = kb; myexit = (pb << 1) | qb; changed = 1; }
                }
              else
                { while (pa != pb && min(pa,pb) < lim)
                    { const bool fa = (pa < pb);
                      const uint32_t pos = fa ? pa : pb;
                      const uint32_t w = win32(D,pos);
                      uint32_t e = mt[w >> 20];
                      if (e == 0u) e = long_entry(t2,symtab,w,&bd);
                      const uint32_t np = pos + DX_E_LEN1(e);
                      if (fa) { pa = np; ca++; } else { pb = np; cb++; }
                    }
                  if (pa == pb) n += cb - ca;
                  else { n = cb; myexit = pb << 1; changed = 1; }
                }
              mystart = want;
            }
          sm.exitst[cur^1][t] = myexit;
          cur ^= 1;
          rounds++;
          if (!__syncthreads_or(changed)) break;
        }
      if (a.dbg != NULL)
        { if (t == 0) { atomicAdd(&a.dbg[symtab*4],(unsigned long long) rounds);
                        atomicAdd(&a.dbg[symtab*4+1],1ull); }
          if (restarts) atomicAdd(&a.dbg[symtab*4+3],(unsigned long long) restarts);
        }

      // ---- place the subsequences ---------------------------------------------------------------
      uint32_t total;
      const uint32_t before = block_excl_scan(n,&total);
      const bool ends_here = (total >= remaining);
      const bool owner = ends_here && before < remaining && remaining <= before + n;
      uint32_t need = 0;
      if (before < remaining) need = min(n,remaining - before);
      if (RUN && symtab == 0 && !owner && need == n && n > 0)
        atomicAdd(&sm.kept_sum,nk);

      // ---- final decode: text, and for the owner of the rlen-th symbol the end of the stream ----
      const bool wr = (dst != NULL);
      const uint32_t outn = min(total,remaining);
      const bool staged = wr && !RUN && outn <= (uint32_t) (kOutStage - 16);
      if (need > 0 && (wr || owner))
        { uint8_t *p;
          if (staged) p = reinterpret_cast<uint8_t *>(sm.u.stage) + before;
          else        p = dst + done + before;
          uint32_t pos = mystart >> 1, cnt = 0, kept = 0, last = 0, bad = 0;
          if (RUN)
            { uint32_t par = mystart & 1u;
              while (cnt < need)
                { if (par == 0)
                    { uint32_t w = win32(D,pos);
                      uint32_t e = rt[w >> 20];
                      if (e == 0u) e = long_single(t2,runtab,w,&bad);
                      uint32_t r = e & 0xffu;
                      last = pos;
                      pos += e >> 8;
                      if (r == 255u) { r = win32(D,pos) >> 16; last = pos; pos += 16; }
                      if (r > need - cnt) { r = need - cnt; bad = 1; }
                      cnt += r; p += r;
                      par = 1;
                      if (cnt >= need) break;
                    }
                  uint32_t w = win32(D,pos);
                  uint32_t e = st[w >> 20];
                  if (e == 0u) e = long_single(t2,symtab,w,&bad);
                  uint32_t c = e & 0xffu;
                  last = pos;
                  pos += e >> 8;
                  if (esc && c == 255u) { c = win32(D,pos) >> 24; last = pos; pos += 8; }
                  if (wr) *p = (uint8_t) c;
                  p++; cnt++;
                  kept += (c != rc);
                  par = 0;
                }
            }
          else
            { uint32_t ppos = pos, pe = 0;
              while (cnt < need)
                { const uint32_t w = win32(D,pos);
                  uint32_t e = mt[w >> 20];
                  if (e == 0u) e = long_entry(t2,symtab,w,&bad);
                  uint32_t c0 = (e >> 16) & 0xffu;
                  if (e & 0x80u) c0 = (w << DX_E_LEN0(e)) >> 24;      // the literal after the escape
                  if (DX_E_N(e) > need - cnt) e = (e & ~0x7fu) | (1u << 5) | DX_E_LEN0(e);
                  if (wr)
                    { p[0] = (uint8_t) c0;
                      if (e & 0x40u) p[1] = (uint8_t) (e >> 24);
                    }
                  ppos = pos; pe = e;
                  pos += DX_E_LEN(e);
                  p   += DX_E_N(e);
                  cnt += DX_E_N(e);
                }
              // position of the last item: the second symbol, or the literal of an escape
              last = ppos + ((pe & 0xc0u) ? DX_E_LEN0(pe) : 0u);
            }
          if (bad) sm.bad = 1;
          if (owner)
            { sm.end_words = (wword*32u + last + 47u) >> 5;   // reference refill rule (QV.c:537-551)
              if (RUN && symtab == 0) atomicAdd(&sm.kept_sum,kept);
            }
        }
      if (staged)
        { __syncthreads();
          const int lane = t & 31, warp = t >> 5;
          for (uint32_t c = (uint32_t) warp*512u; c < outn; c += (kT/32)*512u)
            dx_warp_copy_out(dst + done + c,sm.u.stage + (c >> 2),min(512u,outn - c),lane);
        }
      __syncthreads();
      if (ends_here)
        { words = sm.end_words;
          break;
        }
      done  += total;
      carry  = sm.exitst[cur][nact-1] - ((nact*kS) << 1);
      wword += nact*kSW;
      __syncthreads();
      if (so + (int64_t) wword*4 > a.n + 8)           // ran off the image: corrupt / false start
        { if (t == 0) sm.bad = 1;
          __syncthreads();
          words = wword;
          break;
        }
    }
  if (a.dbg != NULL && t == 0) atomicAdd(&a.dbg[symtab*4+2],1ull);
  if (dst != NULL && t == 0) dst[rlen] = '\n';
  *kept_out = sm.kept_sum;
  __syncthreads();
  return words*4u;
}

// tag line: positions whose deletion QV is the run character get 'n', the others the next packed tag
__device__ void write_tags(const Dec4Args &a, const uint8_t *del, const uint8_t *packed,
                           int32_t rlen, uint8_t *dst)
{ Shared4 &sm = DX_SM;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t caseoff = a.upper ? 32u : 0u;
  uint32_t base_rank = 0;
  for (int32_t p0 = 0; p0 < rlen; p0 += kT*16)
    { const int32_t p = p0 + t*16;
      uint32_t m = 0;
      uint8_t d[16];
      const int cnt = max(0,min(16,rlen - p));
      for (int k = 0; k < cnt; k++)
        { d[k] = del[p+k];
          if (a.delchar < 0 || d[k] != (uint8_t) a.delchar) m |= 1u << k;
        }
      uint32_t tot;
      uint32_t r = base_rank + block_excl_scan(__popc(m),&tot);
      uint32_t wv[4] = { 0, 0, 0, 0 };
      for (int k = 0; k < cnt; k++)
        { uint32_t ch = 'n';
          if (m & (1u << k))
            { const uint32_t byte = packed[r >> 2];
              ch = (0x74676361u >> (8*((byte >> (6 - 2*(r & 3))) & 3u))) & 0xffu;
              r++;
            }
          wv[k >> 2] |= (ch - caseoff) << (8*(k & 3));
        }
      uint32_t *st = sm.u.tagstage[warp];
      st[4*lane] = wv[0]; st[4*lane+1] = wv[1]; st[4*lane+2] = wv[2]; st[4*lane+3] = wv[3];
      __syncwarp();
      const int32_t wbase = p0 + warp*512;
      if (wbase < rlen)
        dx_warp_copy_out(dst + wbase,st,(uint32_t) min(512,rlen - wbase),lane);
      __syncwarp();
      base_rank += tot;
    }
  if (t == 0) dst[rlen] = '\n';
}

__device__ int fmt_int4(uint8_t *p, int32_t v)
{ char tmp[12];
  int  k = 0, len = 0;
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  if (v < 0) p[len++] = '-';
  do { tmp[k++] = (char) ('0' + u % 10); u /= 10; } while (u);
  while (k) p[len++] = (uint8_t) tmp[--k];
  return len;
}

__global__ void __launch_bounds__(kT)
k_qv_decode4(Dec4Args a)
{ Shared4 &sm = DX_SM;
  const int t = threadIdx.x;
  while (true)
    { if (t == 0)
        { sm.entry = (int64_t) atomicAdd(a.ticket,1ull);
          sm.bad = 0; sm.end_words = 0;
        }
      __syncthreads();
      const int64_t e = sm.entry;
      if (e >= a.count) break;
      const int32_t L = a.rlen[e];
      int64_t at = a.start[e];
      int64_t o[6];
      uint8_t *line = NULL;
      if (a.write)
        { const QvDecEntry en = a.ent[e];
          line = a.out + en.text_off;
          if (t == 0)
            { uint8_t *h = a.out + en.out_off;          // "%s/%d/%d_%d RQ=0.%d\n" (undexqv.c:182)
              int hl = 0;
              for (int k = 0; k < a.plen; k++) h[hl++] = (uint8_t) a.prefix[k];
              h[hl++] = '/'; hl += fmt_int4(h+hl,en.well);
              h[hl++] = '/'; hl += fmt_int4(h+hl,en.beg);
              h[hl++] = '_'; hl += fmt_int4(h+hl,en.end);
              const char *rq = " RQ=0.";
              for (int k = 0; k < 6; k++) h[hl++] = (uint8_t) rq[k];
              hl += fmt_int4(h+hl,en.qv);
              h[hl++] = '\n';
            }
        }
      const int64_t stride = (int64_t) L + 1;
      uint32_t kept = 0, dummy;

      o[0] = at;
      if (a.delchar >= 0) at += decode_stream<true >(a,at,L,0,1,a.delchar,line,&kept);
      else                at += decode_stream<false>(a,at,L,0,1,-1,line,&kept);
      o[1] = at;
      const uint32_t clen = (a.delchar < 0) ? (uint32_t) L : kept;
      if (a.write && at + (int64_t) ((clen + 3) >> 2) <= a.n)
        { __syncthreads();                                   // the del line is complete in global memory
          __threadfence_block();
          write_tags(a,line,a.in + at,L,line + stride);
        }
      at += (clen + 3) >> 2;
      o[2] = at;
      at += decode_stream<false>(a,at,L,2,0,-1,a.write ? line + 2*stride : NULL,&dummy);
      o[3] = at;
      at += decode_stream<false>(a,at,L,3,0,-1,a.write ? line + 3*stride : NULL,&dummy);
      o[4] = at;
      if (a.subchar >= 0) at += decode_stream<true >(a,at,L,4,5,a.subchar,a.write ? line + 4*stride : NULL,&dummy);
      else                at += decode_stream<false>(a,at,L,4,5,-1,a.write ? line + 4*stride : NULL,&dummy);
      o[5] = at;
      __syncthreads();
      if (t == 0)
        { const int bad = (sm.bad != 0) || (at > a.n);
          if (a.soff != NULL)
            for (int k = 0; k < 6; k++) a.soff[e*6 + k] = o[k];
          if (a.write) { if (bad) atomicExch(a.status,1); }
          else a.status[e] = bad;
        }
      __syncthreads();
    }
}

}  // namespace

int dxk_qv_decode4(dx_ctx *ctx, const uint8_t *d_in, size_t n, const QvDecTables4 *d_tab,
                   int delchar, int subchar, int upper, int write, int64_t count,
                   const int64_t *d_start, const int32_t *d_rlen, const QvDecEntry *d_ent,
                   const char *d_prefix, int plen, uint8_t *d_out, int64_t *d_soff, int32_t *d_status)
{ if (count == 0) return DX_OK;
  unsigned long long *d_ticket = (unsigned long long *) dx_arena_get(ctx,8);
  if (d_ticket == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_ticket,0,8,ctx->stream));
  Dec4Args a;
  a.in = d_in; a.n = (int64_t) n; a.tab = d_tab;
  a.delchar = delchar; a.subchar = subchar; a.upper = upper; a.write = write;
  a.count = count; a.start = d_start; a.rlen = d_rlen; a.ent = d_ent;
  a.prefix = d_prefix; a.plen = plen; a.out = d_out; a.soff = d_soff; a.status = d_status;
  a.ticket = d_ticket;
  a.dbg = NULL;
  if (getenv("DEXB200_DEBUG") != NULL)
    { a.dbg = (unsigned long long *) dx_arena_get(ctx,32*8);
      if (a.dbg == NULL) return DX_E_NOMEM;
      DX_CUDA(ctx,cudaMemsetAsync(a.dbg,0,32*8,ctx->stream));
    }
  const size_t smem = sizeof(Shared4);
  DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_decode4,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem));
  int64_t grid = (int64_t) ctx->sm_count * 4;
  if (grid > count) grid = count;
  DX_PROF_BEGIN(ctx); k_qv_decode4<<<(unsigned) grid,kT,smem,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,write ? "k_qv_decode4" : "k_qv_walk4");
  if (a.dbg != NULL)
    { unsigned long long h[32];
      DX_CUDA(ctx,cudaMemcpyAsync(h,a.dbg,sizeof(h),cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
      for (int k = 0; k < 5; k++)
        if (h[k*4+2])
          fprintf(stderr,"[dexb200 debug] v4 table %d: streams %llu windows/stream %.2f rounds/window %.2f "
                         "restarts/window %.1f\n",
                  k,h[k*4+2],(double) h[k*4+1]/h[k*4+2],(double) h[k*4]/(h[k*4+1] ? h[k*4+1] : 1),
                  (double) h[k*4+3]/(h[k*4+1] ? h[k*4+1] : 1));
    }
  return DX_OK;
}
