// dx_qv_decode.cu -- .dexqv entry decoding on the device.
//
// Replaces Decode / Decode_Run (reference QV.c:510-691), Packed_Length / Unpack_Tag
// (QV.c:823-847), Uncompress_Read + Lower_Read on the tag block (QV.c:1444-1461) and the
// per-entry text output of undexqv.c:182-207.
//
// A .dexqv stores no lengths: a stream ends where its rlen-th symbol ends, and the next stream
// starts on the following whole 32-bit word (at an arbitrary BYTE address).
//   k_qv_walk    one thread per (candidate) entry decodes lengths only and reports where each of
//                the five streams starts and where the entry ends
//   k_qv_decode  one thread per (entry, stream) decodes symbols into the output text through a
//                16-byte register buffer (aligned 128-bit stores)

#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

// ---- bit reader with the reference's refill rule ------------------------------------------------
// The reference keeps a 16-bit look-ahead window and fetches the next whole word exactly when
// the window would run past the words fetched so far (GET macro, QV.c:537-551).  Counting the
// words fetched therefore gives the stream's length in the file.
struct BitReader
{ const uint8_t *buf;       // image start
  int64_t  n;               // image bytes
  int64_t  at;              // byte offset of the next word to fetch
  uint64_t acc;             // unread bits, left aligned
  int32_t  avail;           // number of valid bits in acc
  int32_t  flip;
  int32_t  bad;

  __device__ __forceinline__ void open(const uint8_t *b, int64_t nbytes, int64_t start, int fl)
  { buf = b; n = nbytes; at = start; acc = 0; avail = 0; flip = fl; bad = 0; }

  __device__ __forceinline__ uint32_t fetch()
  { uint32_t w;
    if (at + 4 > n) { bad = 1; at += 4; return 0; }
    const uint8_t *p = buf + at;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t sh = (uint32_t) (a & 3) * 8;
    const uint32_t *al = reinterpret_cast<const uint32_t *>(a - (a & 3));
    if (sh == 0)
      w = __ldg(al);
    else if (at + 8 <= n)
      w = __funnelshift_r(__ldg(al),__ldg(al+1),sh);
    else
      w = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
    at += 4;
    return flip ? __byte_perm(w,0,0x0123) : w;
  }

  // consume nbits, then make sure 16 bits are visible
  __device__ __forceinline__ void advance(int nbits)
  { acc <<= nbits;
    avail -= nbits;
    if (avail < 16)
      { acc |= (uint64_t) fetch() << (32 - avail);
        avail += 32;
      }
  }
  __device__ __forceinline__ uint32_t window() const { return (uint32_t) (acc >> 48); }
};

struct DecArgs
{ const uint8_t *in;
  int64_t        n;
  const QvDecTables *tab;
  int32_t        delchar, subchar, flip, upper;
};

// Decode (or just walk) one stream of rlen symbols starting at byte `start`; returns the byte
// offset just past it.  SINK::sym(c) / SINK::fill(c,count) receive the output.
template <class SINK>
__device__ int64_t run_stream(const DecArgs &a, int64_t start, int32_t rlen, int symtab, int runtab,
                              int32_t rchar, SINK &sink, int32_t &bad)
{ BitReader br; br.open(a.in,a.n,start,a.flip);
  const uint8_t *slook = a.tab->look[symtab], *slens = a.tab->lens[symtab];
  const int32_t signal = (a.tab->type[symtab] == 2) ? 255 : 256;
  int32_t nb = 0;
  if (rchar < 0)
    { for (int32_t j = 0; j < rlen; j++)
        { br.advance(nb);
          int32_t c = __ldg(slook + br.window());
          nb = __ldg(slens + c);
          if (c == signal)
            { br.advance(nb);
              c = (int32_t) (br.window() >> 8);
              nb = 8;
            }
          sink.sym(c);
        }
    }
  else
    { const uint8_t *rlook = a.tab->look[runtab], *rlens = a.tab->lens[runtab];
      int32_t j = 0;
      while (j < rlen)
        { br.advance(nb);
          int32_t c = __ldg(rlook + br.window());
          nb = __ldg(rlens + c);
          if (c == 255)
            { br.advance(nb);
              c = (int32_t) br.window();
              nb = 16;
            }
          if (c > rlen - j) { bad = 1; c = rlen - j; }       // corrupt run: stay inside the line
          sink.fill(rchar,c);
          j += c;
          if (j < rlen)
            { br.advance(nb);
              c = __ldg(slook + br.window());
              nb = __ldg(slens + c);
              if (c == signal)
                { br.advance(nb);
                  c = (int32_t) (br.window() >> 8);
                  nb = 8;
                }
              sink.sym(c);
              j += 1;
            }
        }
    }
  if (br.bad) bad = 1;
  return br.at;
}

struct CountSink                       // lengths only; counts symbols != delchar (Packed_Length)
{ int32_t rc; int32_t kept;
  __device__ __forceinline__ void sym(int32_t c)              { kept += (c != rc); }
  __device__ __forceinline__ void fill(int32_t, int32_t)      { }
};

struct NullSink
{ __device__ __forceinline__ void sym(int32_t)                { }
  __device__ __forceinline__ void fill(int32_t, int32_t)      { }
};

// ---- walk ---------------------------------------------------------------------------------------

__global__ void k_qv_walk(DecArgs a, const int64_t *start, const int32_t *rlen, int64_t count,
                          int64_t *soff /*[count][6]: 5 stream starts + entry end*/, int32_t *status)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int32_t L = rlen[i];
  int64_t at = start[i];
  int32_t bad = 0;
  int64_t *o = soff + i*6;

  o[0] = at;
  CountSink cs; cs.rc = a.delchar; cs.kept = 0;
  at = run_stream(a,at,L,0,1,a.delchar,cs,bad);
  const int32_t clen = (a.delchar < 0) ? L : cs.kept;
  o[1] = at;
  at += (clen + 3) >> 2;
  NullSink ns;
  o[2] = at;
  at = run_stream(a,at,L,2,0,-1,ns,bad);
  o[3] = at;
  at = run_stream(a,at,L,3,0,-1,ns,bad);
  o[4] = at;
  at = run_stream(a,at,L,4,5,a.subchar,ns,bad);
  o[5] = at;
  if (at > a.n) bad = 1;
  status[i] = bad;
}

// ---- decode -------------------------------------------------------------------------------------

// bytes -> global memory through a 16-byte shift register: byte stores up to the first 16-byte
// boundary, then aligned 128-bit stores, byte stores for the tail
struct TextSink
{ uint8_t *p;               // next byte to write
  uint32_t w0, w1, w2, w3;  // newest byte enters at the top of w3
  int32_t  held;            // bytes held (only after alignment)
  int32_t  head;            // bytes still to be written singly to reach alignment

  __device__ __forceinline__ void open(uint8_t *dst)
  { p = dst; held = 0; w0 = w1 = w2 = w3 = 0;
    head = (int32_t) ((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
  }
  __device__ __forceinline__ void put(uint32_t b)
  { if (head > 0) { *p++ = (uint8_t) b; head--; return; }
    w0 = __funnelshift_r(w0,w1,8); w1 = __funnelshift_r(w1,w2,8);
    w2 = __funnelshift_r(w2,w3,8); w3 = (w3 >> 8) | (b << 24);
    if (++held == 16)
      { dx_stg16(p,make_uint4(w0,w1,w2,w3));
        p += 16; held = 0;
      }
  }
  __device__ __forceinline__ void sym(int32_t c) { put((uint32_t) c & 0xffu); }
  __device__ __forceinline__ void fill_n(uint32_t b, int32_t cnt)
  { while (cnt > 0 && (head > 0 || held != 0)) { put(b); cnt--; }
    if (cnt >= 16)
      { const uint32_t q = b * 0x01010101u;
        const uint4 v = make_uint4(q,q,q,q);
        while (cnt >= 16) { dx_stg16(p,v); p += 16; cnt -= 16; }
      }
    while (cnt > 0) { put(b); cnt--; }
  }
  __device__ __forceinline__ void fill(int32_t c, int32_t cnt) { fill_n((uint32_t) c & 0xffu,cnt); }
  __device__ __forceinline__ void close()
  { // the `held` bytes sit in the top of the register file: oldest first
    for (int32_t k = held; k > 0; k--)
      { uint32_t idx = 16 - k;                       // byte index inside w0..w3
        uint32_t w = (idx & 8) ? ((idx & 4) ? w3 : w2) : ((idx & 4) ? w1 : w0);
        *p++ = (uint8_t) (w >> ((idx & 3)*8));
      }
    held = 0;
  }
};

// tag line: walk the deletion stream again to know which positions kept their tag
struct TagSink
{ TextSink out;
  const uint8_t *tags;      // packed 2-bit tags of this entry
  int32_t  k;               // next packed tag index
  int32_t  rc;
  uint32_t caseoff;         // 0 lower, 32 upper (undexqv.c:198-204 subtracts 32 from every tag)
  __device__ __forceinline__ uint32_t next_tag()
  { const uint32_t b = tags[k >> 2];
    const uint32_t t = (b >> (6 - 2*(k & 3))) & 3u;
    k++;
    return (0x74676361u >> (8*t)) & 0xffu;          // "acgt"
  }
  __device__ __forceinline__ void sym(int32_t c)
  { out.put(((c == rc) ? (uint32_t) 'n' : next_tag()) - caseoff); }
  __device__ __forceinline__ void fill(int32_t, int32_t cnt) { out.fill_n((uint32_t) 'n' - caseoff,cnt); }
};

__device__ int put_int(uint8_t *p, int32_t v)
{ char tmp[12];
  int  k = 0, len = 0;
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  if (v < 0) p[len++] = '-';
  do { tmp[k++] = (char) ('0' + u % 10); u /= 10; } while (u);
  while (k) p[len++] = (uint8_t) tmp[--k];
  return len;
}

__global__ void k_qv_decode(DecArgs a, const QvDecEntry *ent, const int64_t *soff, int64_t count,
                            const char *prefix, int32_t plen, uint8_t *out, int32_t *status)
{ const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count*5) return;
  const int64_t i = t / 5;
  const int     s = (int) (t - i*5);
  const QvDecEntry en = ent[i];
  const int32_t L = en.end - en.beg;
  const int64_t *o = soff + i*6;
  int32_t bad = 0;

  // header text "%s/%d/%d_%d RQ=0.%d\n" (undexqv.c:182)
  uint8_t *h = out + en.out_off;
  int hl = plen;
  if (s == 0)
    { for (int k = 0; k < plen; k++) h[k] = (uint8_t) prefix[k];
      h[hl++] = '/'; hl += put_int(h+hl,en.well);
      h[hl++] = '/'; hl += put_int(h+hl,en.beg);
      h[hl++] = '_'; hl += put_int(h+hl,en.end);
      const char *rq = " RQ=0.";
      for (int k = 0; k < 6; k++) h[hl++] = (uint8_t) rq[k];
      hl += put_int(h+hl,en.qv);
      h[hl++] = '\n';
    }
  uint8_t *line = out + en.text_off + (int64_t) s*((int64_t) L + 1);

  if (s == 1)
    { TagSink ts;
      ts.out.open(line); ts.tags = a.in + o[1]; ts.k = 0; ts.rc = a.delchar;
      ts.caseoff = a.upper ? 32u : 0u;
      if (a.delchar < 0)
        { for (int32_t k = 0; k < L; k++) ts.out.put(ts.next_tag() - ts.caseoff); }
      else
        run_stream(a,o[0],L,0,1,a.delchar,ts,bad);
      ts.out.put('\n');
      ts.out.close();
    }
  else
    { TextSink sk; sk.open(line);
      if (s == 0)      run_stream(a,o[0],L,0,1,a.delchar,sk,bad);
      else if (s == 2) run_stream(a,o[2],L,2,0,-1,sk,bad);
      else if (s == 3) run_stream(a,o[3],L,3,0,-1,sk,bad);
      else             run_stream(a,o[4],L,4,5,a.subchar,sk,bad);
      sk.put('\n');
      sk.close();
    }
  if (bad) atomicExch(status,1);
}

}  // namespace

int dxk_qv_walk(dx_ctx *ctx, const uint8_t *d_in, size_t n, const QvDecTables *d_tab,
                int delchar, int subchar, int flip, const int64_t *d_start, const int32_t *d_rlen,
                int64_t count, int64_t *d_soff, int32_t *d_status)
{ if (count == 0) return DX_OK;
  DecArgs a = { d_in, (int64_t) n, d_tab, delchar, subchar, flip, 0 };
  DX_PROF_BEGIN(ctx); k_qv_walk<<<(unsigned) ((count+63)/64),64,0,ctx->stream>>>(a,d_start,d_rlen,count,d_soff,d_status);
  DX_LAUNCHED(ctx,"k_qv_walk");
  return DX_OK;
}

int dxk_qv_decode(dx_ctx *ctx, const uint8_t *d_in, size_t n, const QvDecTables *d_tab,
                  int delchar, int subchar, int flip, int upper, const QvDecEntry *d_ent,
                  const int64_t *d_soff, int64_t count, const char *d_prefix, int plen,
                  uint8_t *d_out, int32_t *d_status)
{ if (count == 0) return DX_OK;
  DecArgs a = { d_in, (int64_t) n, d_tab, delchar, subchar, flip, upper };
  const int64_t threads = count*5;
  DX_PROF_BEGIN(ctx); k_qv_decode<<<(unsigned) ((threads+63)/64),64,0,ctx->stream>>>(a,d_ent,d_soff,count,d_prefix,plen,
                                                                 d_out,d_status);
  DX_LAUNCHED(ctx,"k_qv_decode");
  return DX_OK;
}
