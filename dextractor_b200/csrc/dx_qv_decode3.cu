// dx_qv_decode3.cu -- parallel .dexqv entry decoder, third generation.
//
// Replaces Decode / Decode_Run (reference QV.c:510-691) + Packed_Length / Unpack_Tag
// (QV.c:823-847) + the per-entry text output of undexqv.c:182-207.
//
// One CTA per entry, every stream decoded by all threads at once.  A Huffman stream cannot be cut
// at known code boundaries, so it is cut into fixed 256-bit subsequences that are decoded
// speculatively (self-synchronising prefix codes):
//
//   0. the window's words (256 subsequences = 8 KB) are staged once in shared memory, already
//      shifted to the stream's byte alignment, so the bit readers are plain shared-memory loads;
//   1. thread i decodes from bit 256*i (a guess) across its subsequence and records, at every
//      32-bit mark, the decoder state it first reaches past that mark (bit position + which table
//      comes next) and the symbols counted so far -- the CHECKPOINTS of its path;
//   2. rounds: a thread whose start differs from its predecessor's exit restarts from that exit,
//      but only until its new path reaches a checkpoint of its old path: from there on both paths
//      are the same, so exit and counts follow by arithmetic.  Thread 0 starts at the true
//      position, so at the fix point every start is a true code boundary (induction over i);
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
constexpr int kSeg      = 32;                   // bits between checkpoints
constexpr int kNSeg     = kS / kSeg;
constexpr int kWinBits  = kT * kS;
constexpr int kWinWords = kWinBits / 32;
constexpr int kPadWords = 8;                    // look-ahead of the last subsequence
constexpr int kOutStage = 3 * kNSeg * kT * 4;   // bytes of the checkpoint arrays, reused as stage

struct Dec3Args
{ const uint8_t *in;
  int64_t        n;
  const QvDecTables2 *tab;
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

struct Shared3
{ uint32_t bits[kWinWords + kPadWords];     // the window's stream words
  uint32_t cp[3][kNSeg][kT];                // checkpoints: state, symbols, kept ; later the out stage
  uint32_t exitst[2][kT];                   // exit state of every subsequence (double buffered)
  uint32_t wsum[kT/32];
  uint32_t end_words, bad, kept_sum;
  int64_t  entry;
  uint32_t tagstage[kT/32][132];
};

// ---- bit reader over the staged words (bit 0 = MSB of word 0) ---------------------------------
struct SBits
{ const uint32_t *w;
  uint32_t idx;
  uint64_t acc;           // unread bits, left aligned
  int32_t  avail;
  __device__ __forceinline__ void seek(const uint32_t *words, uint32_t bit)
  { w = words; idx = bit >> 5;
    const uint32_t w0 = w[idx], w1 = w[idx+1];
    idx += 2;
    const uint32_t off = bit & 31u;
    acc = (((uint64_t) w0 << 32) | w1) << off;
    avail = 64 - (int32_t) off;
  }
  __device__ __forceinline__ void skip(uint32_t nbits)
  { acc <<= nbits;
    avail -= (int32_t) nbits;
    if (avail <= 32)
      { acc |= (uint64_t) w[idx] << (32 - avail);
        idx += 1;
        avail += 32;
      }
  }
  __device__ __forceinline__ uint32_t peek16() const { return (uint32_t) (acc >> 48); }
};

// one table lookup: returns sym | len << 8 ; len 0 only for patterns no code maps to
__device__ __forceinline__ uint32_t lookup(const QvDecTables2 *t, int k, uint32_t w16)
{ uint32_t e = __ldg(&t->prim[k][w16 >> 5]);
  if (e & 0x8000u)
    e = __ldg(&t->sub[k][(e & 0x7fffu)*32u + (w16 & 31u)]);
  return e;
}

struct St { uint32_t pos, par, cnt, kept; };

struct NoSink
{ __device__ __forceinline__ void sym(uint32_t) { }
  __device__ __forceinline__ void fill(uint32_t) { }
};

// byte sink through a generic pointer (shared-memory stage or global line); runs are skipped
// because run-length lines are pre-filled with the run character
struct ByteSink
{ uint8_t *p;
  __device__ __forceinline__ void sym(uint32_t c) { *p++ = (uint8_t) c; }
  __device__ __forceinline__ void fill(uint32_t n) { p += n; }
};

// Decode items from state s until s.pos >= lim or s.cnt >= need.  last_item = bit position of the
// last item read (the literal if the item was escaped).
template <bool RUN, class SINK>
__device__ __forceinline__ void items_until(const QvDecTables2 *tab, int symtab, int runtab,
                                            uint32_t rc, bool esc, SBits &b, St &s, uint32_t lim,
                                            uint32_t need, SINK &sink, uint32_t &last_item,
                                            uint32_t &bad)
{ while (s.pos < lim && s.cnt < need)
    { if (RUN && s.par == 0)
        { const uint32_t e = lookup(tab,runtab,b.peek16());
          uint32_t len = (e >> 8) & 31u, r = e & 0xffu;
          if (len == 0) { len = 1; bad = 1; }
          last_item = s.pos;
          b.skip(len); s.pos += len;
          if (r == 255u)
            { r = b.peek16();
              last_item = s.pos;
              b.skip(16); s.pos += 16;
            }
          if (r > need - s.cnt) { r = need - s.cnt; bad = 1; }
          sink.fill(r);
          s.cnt += r;
          s.par = 1;
          continue;
        }
      const uint32_t e = lookup(tab,symtab,b.peek16());
      uint32_t len = (e >> 8) & 31u, c = e & 0xffu;
      if (len == 0) { len = 1; bad = 1; }
      last_item = s.pos;
      b.skip(len); s.pos += len;
      if (esc && c == 255u)
        { c = b.peek16() >> 8;
          last_item = s.pos;
          b.skip(8); s.pos += 8;
        }
      sink.sym(c);
      s.cnt  += 1;
      s.kept += (c != rc);
      s.par = 0;
    }
}

// block-wide exclusive scan of v (one value per thread); returns exclusive prefix, total in *tot
__device__ __forceinline__ uint32_t block_excl_scan(Shared3 &sm, uint32_t v, uint32_t *tot)
{ const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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

// stage the window's words: stream word j of the window = the LE uint32 at byte p + 4j
__device__ __forceinline__ void stage_window(const Dec3Args &a, Shared3 &sm, const uint8_t *p)
{ const uintptr_t A = reinterpret_cast<uintptr_t>(p);
  const uint32_t *al = reinterpret_cast<const uint32_t *>(A & ~(uintptr_t) 3);
  const uint32_t sh = (uint32_t) (A & 3) * 8;
  const int64_t limit = ((int64_t) (reinterpret_cast<uintptr_t>(a.in + a.n) + 3) -
                         (int64_t) (A & ~(uintptr_t) 3)) >> 2;         // aligned words readable
  for (int j = threadIdx.x; j < kWinWords + kPadWords; j += kT)
    { const uint32_t lo = (j < limit) ? __ldg(al + j) : 0u;
      const uint32_t hi = (sh != 0 && j + 1 < limit) ? __ldg(al + j + 1) : 0u;
      sm.bits[j] = __funnelshift_r(lo,hi,sh);
    }
}

// Decode one stream of `rlen` symbols that starts at byte `so`.  Returns the number of bytes the
// stream occupies; *kept_out = symbol items != rc.  When `dst` is not NULL the line is written.
template <bool RUN>
__device__ __noinline__ uint32_t decode_stream(const Dec3Args &a, Shared3 &sm, int64_t so,
                                               int32_t rlen, int symtab, int runtab, int32_t rci,
                                               uint8_t *dst, uint32_t *kept_out)
{ const int t = threadIdx.x;
  *kept_out = 0;
  if (rlen <= 0) return 0;
  const QvDecTables2 *tab = a.tab;
  const bool esc = (tab->type[symtab] == 2);
  const uint32_t rc = (uint32_t) rci;                 // 0xffffffff for plain streams
  const uint32_t base = (uint32_t) t * kS;
  uint32_t done = 0;                                  // symbols placed by earlier windows
  uint32_t carry = 0;                                 // start state of thread 0 (window relative)
  uint32_t wword = 0;                                 // first stream word of the window
  uint32_t words = 0;
  if (t == 0) sm.kept_sum = 0;
  if (RUN && dst != NULL) fill_line(dst,rc,(uint32_t) rlen);

  while (true)
    { stage_window(a,sm,a.in + so + (int64_t) wword*4);
      __syncthreads();

      // ---- round 0: speculative decode of the own subsequence, with checkpoints ---------------
      NoSink ns;
      uint32_t li = 0, bd = 0;
      uint32_t mystart = (t == 0) ? carry : (base << 1);
      uint32_t myexit, n, nk;
      { St s; s.pos = mystart >> 1; s.par = mystart & 1u; s.cnt = 0; s.kept = 0;
        SBits b; b.seek(sm.bits,s.pos);
#pragma unroll 1
        for (int seg = 0; seg < kNSeg; seg++)
          { items_until<RUN>(tab,symtab,runtab,rc,esc,b,s,base + (uint32_t) (seg+1)*kSeg,
                             0xffffffffu,ns,li,bd);
            sm.cp[0][seg][t] = (s.pos << 1) | s.par;
            sm.cp[1][seg][t] = s.cnt;
            sm.cp[2][seg][t] = s.kept;
          }
        myexit = (s.pos << 1) | s.par; n = s.cnt; nk = s.kept;
      }
      int cur = 0;
      sm.exitst[0][t] = myexit;
      __syncthreads();

      // ---- rounds: adopt the predecessor's exit, re-decode until the old path is met ----------
      uint32_t rounds = 0, restarts = 0;
      while (true)
        { const uint32_t want = (t > 0) ? sm.exitst[cur][t-1] : mystart;
          int changed = 0;
          if (want != mystart)
            { mystart = want;
              restarts++;
              St s; s.pos = want >> 1; s.par = want & 1u; s.cnt = 0; s.kept = 0;
              SBits b; b.seek(sm.bits,s.pos);
              bool merged = false;
              uint32_t st = want;
#pragma unroll 1
              for (int seg = (int) ((s.pos - base) >> 5); seg < kNSeg; seg++)
                { items_until<RUN>(tab,symtab,runtab,rc,esc,b,s,base + (uint32_t) (seg+1)*kSeg,
                                   0xffffffffu,ns,li,bd);
                  st = (s.pos << 1) | s.par;
                  if (st == sm.cp[0][seg][t])
                    { const uint32_t dn = s.cnt  - sm.cp[1][seg][t];
                      const uint32_t dk = s.kept - sm.cp[2][seg][t];
                      for (int k = seg; k < kNSeg; k++)
                        { sm.cp[1][k][t] += dn; sm.cp[2][k][t] += dk; }
                      n += dn; nk += dk;
                      merged = true;
                      break;
                    }
                  sm.cp[0][seg][t] = st;
                  sm.cp[1][seg][t] = s.cnt;
                  sm.cp[2][seg][t] = s.kept;
                }
              if (!merged)
                { n = s.cnt; nk = s.kept;
                  if (st != myexit) { myexit = st; changed = 1; }
                }
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
      const uint32_t before = block_excl_scan(sm,n,&total);
      const uint32_t remaining = (uint32_t) rlen - done;
      const bool ends_here = (total >= remaining);
      const bool owner = ends_here && before < remaining && remaining <= before + n;
      uint32_t need = 0;
      if (before < remaining) need = min(n,remaining - before);

      if (owner)
        { // this thread holds the rlen-th symbol: find the exact end of the stream
          St s; s.pos = mystart >> 1; s.par = mystart & 1u; s.cnt = 0; s.kept = 0;
          SBits b; b.seek(sm.bits,s.pos);
          uint32_t last = 0, bad = 0;
          items_until<RUN>(tab,symtab,runtab,rc,esc,b,s,0xffffffffu,need,ns,last,bad);
          sm.end_words = (wword*32u + last + 47u) >> 5;       // reference refill rule (QV.c:537-551)
          if (bad) sm.bad = 1;
          if (symtab == 0) atomicAdd(&sm.kept_sum,s.kept);
        }
      else if (symtab == 0 && need == n && n > 0)
        atomicAdd(&sm.kept_sum,nk);

      // ---- final decode, producing text ---------------------------------------------------------
      if (dst != NULL)
        { const uint32_t outn = min(total,remaining);
          const bool staged = !RUN && outn <= (uint32_t) (kOutStage - 16);
          if (staged) __syncthreads();                 // everyone is done with the checkpoints
          if (need > 0)
            { ByteSink bs;
              if (RUN)         bs.p = dst + done + before;
              else if (staged) bs.p = reinterpret_cast<uint8_t *>(&sm.cp[0][0][0]) + before;
              else             bs.p = dst + done + before;
              St s; s.pos = mystart >> 1; s.par = mystart & 1u; s.cnt = 0; s.kept = 0;
              SBits b; b.seek(sm.bits,s.pos);
              uint32_t last = 0, bad = 0;
              items_until<RUN>(tab,symtab,runtab,rc,esc,b,s,0xffffffffu,need,bs,last,bad);
              if (bad) sm.bad = 1;
            }
          if (staged)
            { __syncthreads();
              const int lane = t & 31, warp = t >> 5;
              const uint32_t *stw = &sm.cp[0][0][0];
              for (uint32_t c = (uint32_t) warp*512u; c < outn; c += (kT/32)*512u)
                dx_warp_copy_out(dst + done + c,stw + (c >> 2),min(512u,outn - c),lane);
            }
        }
      __syncthreads();
      if (ends_here)
        { words = sm.end_words;
          break;
        }
      done  += total;
      carry  = sm.exitst[cur][kT-1] - ((uint32_t) kWinBits << 1);
      wword += kWinWords;
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
__device__ void write_tags(const Dec3Args &a, Shared3 &sm, const uint8_t *del, const uint8_t *packed,
                           int32_t rlen, uint8_t *dst)
{ const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
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
      uint32_t r = base_rank + block_excl_scan(sm,__popc(m),&tot);
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
      uint32_t *st = sm.tagstage[warp];
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

__device__ int fmt_int3(uint8_t *p, int32_t v)
{ char tmp[12];
  int  k = 0, len = 0;
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  if (v < 0) p[len++] = '-';
  do { tmp[k++] = (char) ('0' + u % 10); u /= 10; } while (u);
  while (k) p[len++] = (uint8_t) tmp[--k];
  return len;
}

__global__ void __launch_bounds__(kT)
k_qv_decode3(Dec3Args a)
{ __shared__ Shared3 sm;
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
              h[hl++] = '/'; hl += fmt_int3(h+hl,en.well);
              h[hl++] = '/'; hl += fmt_int3(h+hl,en.beg);
              h[hl++] = '_'; hl += fmt_int3(h+hl,en.end);
              const char *rq = " RQ=0.";
              for (int k = 0; k < 6; k++) h[hl++] = (uint8_t) rq[k];
              hl += fmt_int3(h+hl,en.qv);
              h[hl++] = '\n';
            }
        }
      const int64_t stride = (int64_t) L + 1;
      uint32_t kept = 0, dummy;

      o[0] = at;
      if (a.delchar >= 0) at += decode_stream<true >(a,sm,at,L,0,1,a.delchar,line,&kept);
      else                at += decode_stream<false>(a,sm,at,L,0,1,-1,line,&kept);
      o[1] = at;
      const uint32_t clen = (a.delchar < 0) ? (uint32_t) L : kept;
      if (a.write && at + (int64_t) ((clen + 3) >> 2) <= a.n)
        { __syncthreads();                                   // the del line is complete in global memory
          __threadfence_block();
          write_tags(a,sm,line,a.in + at,L,line + stride);
        }
      at += (clen + 3) >> 2;
      o[2] = at;
      at += decode_stream<false>(a,sm,at,L,2,0,-1,a.write ? line + 2*stride : NULL,&dummy);
      o[3] = at;
      at += decode_stream<false>(a,sm,at,L,3,0,-1,a.write ? line + 3*stride : NULL,&dummy);
      o[4] = at;
      if (a.subchar >= 0) at += decode_stream<true >(a,sm,at,L,4,5,a.subchar,a.write ? line + 4*stride : NULL,&dummy);
      else                at += decode_stream<false>(a,sm,at,L,4,5,-1,a.write ? line + 4*stride : NULL,&dummy);
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

int dxk_qv_decode3(dx_ctx *ctx, const uint8_t *d_in, size_t n, const QvDecTables2 *d_tab,
                   int delchar, int subchar, int upper, int write, int64_t count,
                   const int64_t *d_start, const int32_t *d_rlen, const QvDecEntry *d_ent,
                   const char *d_prefix, int plen, uint8_t *d_out, int64_t *d_soff, int32_t *d_status)
{ if (count == 0) return DX_OK;
  unsigned long long *d_ticket = (unsigned long long *) dx_arena_get(ctx,8);
  if (d_ticket == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_ticket,0,8,ctx->stream));
  Dec3Args a;
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
  int64_t grid = (int64_t) ctx->sm_count * 5;
  if (grid > count) grid = count;
  DX_PROF_BEGIN(ctx); k_qv_decode3<<<(unsigned) grid,kT,0,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,write ? "k_qv_decode3" : "k_qv_walk3");
  if (a.dbg != NULL)
    { unsigned long long h[32];
      DX_CUDA(ctx,cudaMemcpyAsync(h,a.dbg,sizeof(h),cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
      for (int k = 0; k < 5; k++)
        if (h[k*4+2])
          fprintf(stderr,"[dexb200 debug] v3 table %d: streams %llu windows/stream %.2f rounds/window %.2f "
                         "restarts/window %.1f\n",
                  k,h[k*4+2],(double) h[k*4+1]/h[k*4+2],(double) h[k*4]/(h[k*4+1] ? h[k*4+1] : 1),
                  (double) h[k*4+3]/(h[k*4+1] ? h[k*4+1] : 1));
    }
  return DX_OK;
}
