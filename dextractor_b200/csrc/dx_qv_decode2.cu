// dx_qv_decode2.cu -- parallel .dexqv entry decoder: one CTA per entry, every stream decoded by
// all threads at once.
//
// Replaces Decode / Decode_Run (reference QV.c:510-691) + Packed_Length / Unpack_Tag
// (QV.c:823-847) + the per-entry text output of undexqv.c:182-207, like dx_qv_decode.cu, but
// without the serial chain inside a stream.  A Huffman stream cannot be cut at known code
// boundaries, so the stream is cut into fixed 128-bit subsequences and decoded speculatively:
//
//   1. thread i decodes from bit 128*i (a guess) until it crosses bit 128*(i+1) and publishes
//      where it stopped (exit position + which table comes next) and how many symbols it saw;
//   2. every thread whose start differs from its predecessor's exit restarts from that exit;
//      repeat until nothing changes.  Thread 0 starts at the true position, so at the fix point
//      EVERY start is a true code boundary (induction over i) -- prefix codes resynchronise
//      after a few symbols, so this takes 2-3 rounds instead of one per subsequence;
//   3. an exclusive scan of the symbol counts places every subsequence in the output line and
//      finds the subsequence in which the rlen-th symbol -- hence the stream -- ends; the
//      stream's length in the file follows from the position of its last item ((p_last+47)>>5
//      words, the reference's refill rule, QV.c:537-551);
//   4. every thread decodes its subsequence once more, now writing text.
// Speculation only costs time: nothing is written before the fix point is reached.

#include <stdio.h>
#include <stdlib.h>
#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSubBits = 128;                       // bits per subsequence
constexpr int kWinBits = kThreads * kSubBits;       // bits per window

struct Dec2Args
{ const uint8_t *in;
  int64_t        n;
  const QvDecTables2 *tab;
  int32_t        delchar, subchar, upper, write;
  // entries
  int64_t        count;
  const int64_t *start;        // first stream byte of each entry (after beg/end/qv)
  const int32_t *rlen;
  const QvDecEntry *ent;       // write mode: output placement
  const char    *prefix; int32_t plen;
  uint8_t       *out;
  int64_t       *soff;         // [count][6] or NULL
  int32_t       *status;       // [count] (walk) or [1] (decode)
  unsigned long long *ticket;
  unsigned long long *dbg;     // optional counters: [kind][0 rounds, 1 windows, 2 streams]
};

// ---- bit reader over 32-bit words at an arbitrary byte address --------------------------------
struct Bits
{ const uint32_t *al;     // aligned word pointer of the stream start
  uint32_t sh;            // byte misalignment * 8
  int64_t  limit;         // aligned words readable (beyond: zeros)
  int64_t  idx;           // next aligned word index to load (hi of the next fetch)
  uint32_t lo;
  uint64_t acc;           // unread bits, left aligned
  int32_t  avail;

  __device__ __forceinline__ uint32_t ldw(int64_t k) const
  { return (k < limit) ? __ldg(al + k) : 0u; }

  __device__ __forceinline__ void seek(const uint8_t *stream, const uint8_t *image_end, uint32_t bitpos)
  { const uintptr_t a = reinterpret_cast<uintptr_t>(stream);
    al = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t) 3);
    sh = (uint32_t) (a & 3) * 8;
    limit = (reinterpret_cast<uintptr_t>(image_end) + 3 - (a & ~(uintptr_t) 3)) >> 2;
    idx = bitpos >> 5;
    lo = ldw(idx); idx++;
    const uint32_t w0 = fetch(), w1 = fetch();
    const uint32_t off = bitpos & 31u;
    acc = (((uint64_t) w0 << 32) | w1) << off;
    avail = 64 - (int32_t) off;
  }
  __device__ __forceinline__ uint32_t fetch()
  { const uint32_t hi = ldw(idx); idx++;
    const uint32_t w = __funnelshift_r(lo,hi,sh);
    lo = hi;
    return w;
  }
  __device__ __forceinline__ void skip(uint32_t nbits)
  { acc <<= nbits;
    avail -= (int32_t) nbits;
    if (avail <= 32)
      { acc |= (uint64_t) fetch() << (32 - avail);
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

struct Span               // what a thread learns from decoding one subsequence
{ uint32_t exit_pos;      // first bit not consumed (>= the subsequence's upper limit)
  uint32_t exit_par;      // 1 if a run item was read and its symbol item is still to come
  uint32_t nsym;          // output symbols produced
  uint32_t nkept;         // symbol items different from the run character (tags kept)
};

// SINK: void sym(c), void fill(c, n)
struct NoSink
{ __device__ __forceinline__ void sym(uint32_t) { }
  __device__ __forceinline__ void fill(uint32_t, uint32_t) { }
};

// Decode items starting at (pos,par) until pos >= limit_bit or `need` symbols were produced.
// symtab/runtab: table indices; rc < 0 for plain streams.  last_item returns the bit position of
// the last item read (the literal if the item was escaped).
template <class SINK>
__device__ __forceinline__ Span decode_span(const Dec2Args &a, const uint8_t *stream, int symtab,
                                            int runtab, int32_t rc, uint32_t pos, uint32_t par,
                                            uint32_t limit_bit, uint32_t need, SINK &sink,
                                            uint32_t &last_item, int32_t &bad)
{ Bits b; b.seek(stream,a.in + a.n,pos);
  const bool esc = (a.tab->type[symtab] == 2);
  Span s; s.nsym = 0; s.nkept = 0;
  while (pos < limit_bit && s.nsym < need)
    { if (rc >= 0 && par == 0)
        { uint32_t e = lookup(a.tab,runtab,b.peek16());
          uint32_t len = (e >> 8) & 31u, r = e & 0xffu;
          if (len == 0) { len = 1; bad = 1; }
          last_item = pos;
          b.skip(len); pos += len;
          if (r == 255u)
            { r = b.peek16();
              last_item = pos;
              b.skip(16); pos += 16;
            }
          if (r > need - s.nsym) { r = need - s.nsym; bad = 1; }
          sink.fill((uint32_t) rc,r);
          s.nsym += r;
          par = 1;
          continue;
        }
      uint32_t e = lookup(a.tab,symtab,b.peek16());
      uint32_t len = (e >> 8) & 31u, c = e & 0xffu;
      if (len == 0) { len = 1; bad = 1; }
      last_item = pos;
      b.skip(len); pos += len;
      if (esc && c == 255u)
        { c = b.peek16() >> 8;
          last_item = pos;
          b.skip(8); pos += 8;
        }
      sink.sym(c);
      s.nsym += 1;
      s.nkept += (c != (uint32_t) rc);
      par = 0;
    }
  s.exit_pos = pos; s.exit_par = par;
  return s;
}

// text bytes -> global memory through a 16-byte shift register (see dx_qv_decode.cu)
struct LineSink
{ uint8_t *p;
  uint32_t w0, w1, w2, w3;
  int32_t  held, head;
  __device__ __forceinline__ void open(uint8_t *dst)
  { p = dst; held = 0; w0 = w1 = w2 = w3 = 0;
    head = (int32_t) ((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
  }
  __device__ __forceinline__ void put(uint32_t c)
  { if (head > 0) { *p++ = (uint8_t) c; head--; return; }
    w0 = __funnelshift_r(w0,w1,8); w1 = __funnelshift_r(w1,w2,8);
    w2 = __funnelshift_r(w2,w3,8); w3 = (w3 >> 8) | (c << 24);
    if (++held == 16)
      { dx_stg16(p,make_uint4(w0,w1,w2,w3));
        p += 16; held = 0;
      }
  }
  __device__ __forceinline__ void sym(uint32_t c) { put(c & 0xffu); }
  __device__ __forceinline__ void fill(uint32_t c, uint32_t cnt)
  { while (cnt > 0 && (head > 0 || held != 0)) { put(c); cnt--; }
    if (cnt >= 16)
      { const uint32_t q = c * 0x01010101u;
        const uint4 v = make_uint4(q,q,q,q);
        while (cnt >= 16) { dx_stg16(p,v); p += 16; cnt -= 16; }
      }
    while (cnt > 0) { put(c); cnt--; }
  }
  __device__ __forceinline__ void close()
  { for (int32_t k = held; k > 0; k--)
      { const uint32_t idx = 16 - k;
        const uint32_t w = (idx & 8) ? ((idx & 4) ? w3 : w2) : ((idx & 4) ? w1 : w0);
        *p++ = (uint8_t) (w >> ((idx & 3)*8));
      }
    held = 0;
  }
};

struct Shared
{ uint32_t exit_pos[kThreads];
  uint32_t exit_par[kThreads];
  uint32_t nsym[kThreads];
  uint32_t nkept[kThreads];
  uint32_t scan[kThreads];
  uint32_t wsum[kThreads/32];
  uint32_t total, total_kept, end_words, done, bad;
  uint32_t carry_pos, carry_par;
  int64_t  entry;
  uint32_t tagstage[kThreads/32][132];
};

// block-wide exclusive scan of v (one value per thread); returns exclusive prefix, total in *tot
__device__ __forceinline__ uint32_t block_excl_scan(Shared &sm, uint32_t v, uint32_t *tot)
{ const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = dx_warp_incl_sum(v,lane);
  if (lane == 31) sm.wsum[warp] = inc;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kThreads/32; w++)
    { const uint32_t x = sm.wsum[w];
      if (w < warp) before += x;
      all += x;
    }
  __syncthreads();
  *tot = all;
  return before + inc - v;
}

// Decode one stream of `rlen` symbols that starts at byte `so`.  Returns the number of bytes the
// stream occupies; *kept = symbol items != rc.  When `dst` is not NULL the line is written there.
__device__ uint32_t decode_stream(const Dec2Args &a, Shared &sm, int64_t so, int32_t rlen,
                                  int symtab, int runtab, int32_t rc, uint8_t *dst, uint32_t *kept)
{ const int t = threadIdx.x;
  const uint8_t *stream = a.in + so;
  uint32_t done_syms = 0, kept_total = 0;
  uint32_t win = 0;                                  // window base bit
  uint32_t carry_pos = 0, carry_par = 0;
  uint32_t words = 0;
  *kept = 0;
  if (rlen <= 0) return 0;

  while (true)
    { const uint32_t lim = win + (uint32_t) (t+1)*kSubBits;
      uint32_t my_pos = (t == 0) ? carry_pos : win + (uint32_t) t*kSubBits;
      uint32_t my_par = (t == 0) ? carry_par : 0u;
      uint32_t dummy_last; int32_t dummy_bad = 0;
      NoSink ns;
      // a subsequence that begins beyond the image is empty
      const bool live = (so + (int64_t) (my_pos >> 3) < a.n);
      Span s;
      s.exit_pos = max(my_pos,lim); s.exit_par = my_par; s.nsym = 0; s.nkept = 0;
      if (live && my_pos < lim)
        s = decode_span(a,stream,symtab,runtab,rc,my_pos,my_par,lim,0xffffffffu,ns,dummy_last,dummy_bad);
      sm.exit_pos[t] = s.exit_pos; sm.exit_par[t] = s.exit_par;
      __syncthreads();

      // synchronise starts with predecessors' exits until nothing changes
      while (true)
        { uint32_t want_pos = my_pos, want_par = my_par;
          if (t > 0) { want_pos = sm.exit_pos[t-1]; want_par = sm.exit_par[t-1]; }
          __syncthreads();
          int changed = 0;
          if (t > 0 && (want_pos != my_pos || want_par != my_par))
            { my_pos = want_pos; my_par = want_par;
              s.exit_pos = max(my_pos,lim); s.exit_par = my_par; s.nsym = 0; s.nkept = 0;
              if (my_pos < lim && so + (int64_t) (my_pos >> 3) < a.n)
                s = decode_span(a,stream,symtab,runtab,rc,my_pos,my_par,lim,0xffffffffu,ns,
                                dummy_last,dummy_bad);
              sm.exit_pos[t] = s.exit_pos; sm.exit_par[t] = s.exit_par;
              changed = 1;
            }
          if (a.dbg != NULL && t == 0) atomicAdd(&a.dbg[symtab*4],1ull);
          if (!__syncthreads_or(changed)) break;
        }
      if (a.dbg != NULL && t == 0) atomicAdd(&a.dbg[symtab*4+1],1ull);

      // place the subsequences: exclusive scan of symbol counts
      uint32_t total;
      const uint32_t before = block_excl_scan(sm,s.nsym,&total);
      const uint32_t remaining = (uint32_t) rlen - done_syms;
      const bool ends_here = (total >= remaining);

      if (ends_here && before < remaining && remaining <= before + s.nsym)
        { // this thread holds the rlen-th symbol: find the exact end of the stream
          uint32_t last = 0; int32_t bad = 0;
          Span e = decode_span(a,stream,symtab,runtab,rc,my_pos,my_par,lim + 64,remaining - before,
                               ns,last,bad);
          (void) e;
          sm.end_words = (last + 47u) >> 5;           // reference refill rule (QV.c:537-551)
          if (bad) sm.bad = 1;
        }
      // kept tags (symbol items != rc) up to the end of the stream
      { uint32_t mykept = s.nkept;
        if (ends_here && before + s.nsym > remaining)
          { mykept = 0;
            if (before < remaining)
              { uint32_t last; int32_t bad = 0;
                Span e = decode_span(a,stream,symtab,runtab,rc,my_pos,my_par,lim + 64,
                                     remaining - before,ns,last,bad);
                mykept = e.nkept;
              }
          }
        uint32_t ktot;
        block_excl_scan(sm,mykept,&ktot);
        kept_total += ktot;
      }

      if (dst != NULL && before < remaining && s.nsym > 0)
        { LineSink ls; ls.open(dst + done_syms + before);
          uint32_t last; int32_t bad = 0;
          decode_span(a,stream,symtab,runtab,rc,my_pos,my_par,lim + 64,
                      min(s.nsym,remaining - before),ls,last,bad);
          ls.close();
          if (bad) sm.bad = 1;
        }
      __syncthreads();
      if (ends_here)
        { words = sm.end_words;
          break;
        }
      done_syms += total;
      carry_pos = sm.exit_pos[kThreads-1];
      carry_par = sm.exit_par[kThreads-1];
      win += kWinBits;
      __syncthreads();
      if ((int64_t) so + (win >> 3) > a.n + 8)        // ran off the image: corrupt / false start
        { if (t == 0) sm.bad = 1;
          __syncthreads();
          words = (win >> 5);
          break;
        }
    }
  if (a.dbg != NULL && t == 0) atomicAdd(&a.dbg[symtab*4+2],1ull);
  if (dst != NULL && t == 0) dst[rlen] = '\n';
  *kept = kept_total;
  return words*4u;
}

// tag line: positions whose deletion QV is the run character get 'n', the others the next packed tag
__device__ void write_tags(const Dec2Args &a, Shared &sm, const uint8_t *del, const uint8_t *packed,
                           int32_t rlen, uint8_t *dst)
{ const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t caseoff = a.upper ? 32u : 0u;
  uint32_t base_rank = 0;
  for (int32_t p0 = 0; p0 < rlen; p0 += kThreads*16)
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

__device__ int fmt_int2(uint8_t *p, int32_t v)
{ char tmp[12];
  int  k = 0, len = 0;
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  if (v < 0) p[len++] = '-';
  do { tmp[k++] = (char) ('0' + u % 10); u /= 10; } while (u);
  while (k) p[len++] = (uint8_t) tmp[--k];
  return len;
}

__global__ void __launch_bounds__(kThreads)
k_qv_decode2(Dec2Args a)
{ __shared__ Shared sm;
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
              h[hl++] = '/'; hl += fmt_int2(h+hl,en.well);
              h[hl++] = '/'; hl += fmt_int2(h+hl,en.beg);
              h[hl++] = '_'; hl += fmt_int2(h+hl,en.end);
              const char *rq = " RQ=0.";
              for (int k = 0; k < 6; k++) h[hl++] = (uint8_t) rq[k];
              hl += fmt_int2(h+hl,en.qv);
              h[hl++] = '\n';
            }
        }
      const int64_t stride = (int64_t) L + 1;
      uint32_t kept = 0, dummy;

      o[0] = at;
      at += decode_stream(a,sm,at,L,0,1,a.delchar,line,&kept);
      o[1] = at;
      const uint32_t clen = (a.delchar < 0) ? (uint32_t) L : kept;
      if (a.write && at + (int64_t) ((clen + 3) >> 2) <= a.n)
        { __syncthreads();                                   // the del line is complete in global memory
          __threadfence_block();
          write_tags(a,sm,line,a.in + at,L,line + stride);
        }
      at += (clen + 3) >> 2;
      o[2] = at;
      at += decode_stream(a,sm,at,L,2,0,-1,a.write ? line + 2*stride : NULL,&dummy);
      o[3] = at;
      at += decode_stream(a,sm,at,L,3,0,-1,a.write ? line + 3*stride : NULL,&dummy);
      o[4] = at;
      at += decode_stream(a,sm,at,L,4,5,a.subchar,a.write ? line + 4*stride : NULL,&dummy);
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

int dxk_qv_decode2(dx_ctx *ctx, const uint8_t *d_in, size_t n, const QvDecTables2 *d_tab,
                   int delchar, int subchar, int upper, int write, int64_t count,
                   const int64_t *d_start, const int32_t *d_rlen, const QvDecEntry *d_ent,
                   const char *d_prefix, int plen, uint8_t *d_out, int64_t *d_soff, int32_t *d_status)
{ if (count == 0) return DX_OK;
  unsigned long long *d_ticket = (unsigned long long *) dx_arena_get(ctx,8);
  if (d_ticket == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_ticket,0,8,ctx->stream));
  Dec2Args a;
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
  int64_t grid = (int64_t) ctx->sm_count * 6;
  if (grid > count) grid = count;
  DX_PROF_BEGIN(ctx); k_qv_decode2<<<(unsigned) grid,kThreads,0,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,write ? "k_qv_decode2" : "k_qv_walk2");
  if (a.dbg != NULL)
    { unsigned long long h[32];
      DX_CUDA(ctx,cudaMemcpyAsync(h,a.dbg,sizeof(h),cudaMemcpyDeviceToHost,ctx->stream));
      DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
      for (int k = 0; k < 5; k++)
        if (h[k*4+2])
          fprintf(stderr,"[dexb200 debug] table %d: streams %llu windows/stream %.2f sync rounds/window %.2f\n",
                  k,h[k*4+2],(double) h[k*4+1]/h[k*4+2],(double) h[k*4]/(h[k*4+1] ? h[k*4+1] : 1));
    }
  return DX_OK;
}
