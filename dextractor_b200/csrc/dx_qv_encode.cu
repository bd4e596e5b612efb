// dx_qv_encode.cu -- pass 2 of the QV coder on the device.
//
// Replaces Encode / Encode_Run (reference QV.c:386-506), Pack_Tag + Number_Read + Compress_Read
// on the tag line (QV.c:810-819, 1400-1404), the lossy masks (QV.c:1406-1415) and the per-entry
// header writes of dexqv.c:128-139.
//
//   k_qv_size     one warp per (entry, stream): bits of every item, position of the last item
//                 -> 32-bit words the reference would write (the (p_last+47)>>5 rule,
//                 QV.c:436-442), kept-tag count -> tag bytes
//   k_qv_offsets  exclusive scan of the per-entry byte totals (the implicit file position of
//                 the reference's fwrite calls)
//   k_qv_emit     one warp per (entry, stream): every lane codes 16 consecutive symbols, a warp
//                 scan gives each lane its bit offset, lanes OR their codes MSB-first into a
//                 shared-memory staging area, which is flushed with aligned 32-bit stores at
//                 whatever byte alignment the stream has in the file

#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kEncWarps   = 16;
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kStageWords = 1024;                 // per warp; a 512-symbol row needs <= 897
constexpr int kFetch      = 4;

struct EncArgs
{ const uint8_t  *text;
  const uint8_t  *text_end16;     // first address past the readable 16-byte-padded text
  QvEntries       ent;
  const uint32_t *tab;            // [6][256] packed entries (DX_ENC_*)
  int32_t         delchar, subchar, lossy, lwell_in;
  uint32_t       *bytes;          // [n][6]
  int64_t        *off;            // [n+1]
  uint8_t        *out;
  unsigned long long *ticket;
};

__device__ __forceinline__ uint32_t item_len(uint32_t e, uint32_t lit)
{ return DX_ENC_LEN(e) + (DX_ENC_ESC(e) ? lit : 0u); }

// local bit accumulator: whole words go to the staging area with one atomicOr each
struct BitSink
{ uint32_t *stage; uint32_t w, cur, fill;
  __device__ __forceinline__ void start(uint32_t *st, uint32_t pos)
  { stage = st; w = pos >> 5; fill = pos & 31u; cur = 0; }
  __device__ __forceinline__ void put(uint32_t code, uint32_t len)
  { if (fill + len < 32u)
      { cur |= code << (32u - fill - len); fill += len; }
    else
      { const uint32_t spill = fill + len - 32u;
        cur |= code >> spill;
        atomicOr(&stage[w],cur);
        w += 1;
        cur  = spill ? (code << (32u - spill)) : 0u;
        fill = spill;
      }
  }
  __device__ __forceinline__ void done() { if (cur) atomicOr(&stage[w],cur); }
};

__device__ __forceinline__ uint32_t base2(uint32_t c)        // Number_Read, DB.c:394-411
{ c |= 0x20u;
  return (c == 'c') ? 1u : (c == 'g') ? 2u : (c == 't') ? 3u : 0u;
}

// Everything one warp needs to walk one line in rows of 32 chunks x 16 bytes.
struct LineWalk
{ const uint8_t *base;      // 16-byte aligned address at or before the line
  int32_t skew, rlen, nchunk;
  __device__ __forceinline__ void set(const uint8_t *line, int32_t len)
  { skew = (int32_t) (reinterpret_cast<uintptr_t>(line) & 15);
    base = line - skew; rlen = len; nchunk = (skew + len + 15) >> 4;
  }
  __device__ __forceinline__ uint32_t valid(int32_t c) const
  { if (c >= nchunk) return 0;
    const int32_t p0 = c*16 - skew;
    return dx_range16(max(0,-p0),min(16,rlen - p0));
  }
};

// MODE 0: size only.  MODE 1: size + emit.
// Returns (through refs) the stream's total bits and the bit position of its last item.
// In MODE 1 the staging area holds the not yet flushed tail; `gptr` advances over flushed bytes.
template <int MODE>
__device__ void code_stream(const EncArgs &a, const uint32_t *stab, const uint8_t *line,
                            int32_t rlen, int kind /*0 del 2 ins 3 mrg 4 sub*/, int lane,
                            uint32_t *stage, uint8_t *&gptr,
                            uint32_t &total_bits, uint32_t &plast_out)
{ const int32_t rc = (kind == 0) ? a.delchar : (kind == 4) ? a.subchar : -1;
  const uint32_t *sym = stab + kind*256;
  const uint32_t *run = stab + (kind == 0 ? 1 : 5)*256;
  const uint32_t lossmask = !a.lossy ? 0xffffffffu : (kind == 2) ? 0xfefefefeu
                                                    : (kind == 3) ? 0xfcfcfcfcu : 0xffffffffu;
  LineWalk lw; lw.set(line,rlen);

  uint32_t flushed_bits = 0;      // bits already written to global (multiple of 32)
  uint32_t stage_bits   = 0;      // bits currently staged (from word 0 of stage)
  int32_t  prev  = -1;            // last non-run position so far
  int32_t  plast = -1;            // lane-local candidate for the last item's bit position

  for (int32_t c0 = 0; c0 < lw.nchunk; c0 += 128)
    { uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
        { const int32_t c = c0 + j*32 + lane;
          v[j] = (c < lw.nchunk) ? dx_ldg16(lw.base + (int64_t) c*16) : make_uint4(0,0,0,0);
          v[j].x &= lossmask; v[j].y &= lossmask; v[j].z &= lossmask; v[j].w &= lossmask;
        }
#pragma unroll
      for (int j = 0; j < 4; j++)
        { const int32_t c  = c0 + j*32 + lane;
          if (c0 + j*32 >= lw.nchunk) break;                        // warp-uniform
          const int32_t p0 = c*16 - lw.skew;
          const uint32_t valid = lw.valid(c);
          uint32_t lane_bits = 0;
          uint32_t m = valid;
          int32_t  pv = -1;

          if (rc < 0)
            { uint32_t mm = m;
              while (mm)
                { const int i = __ffs(mm) - 1; mm &= mm - 1;
                  lane_bits += item_len(sym[dx_byte_of(v[j],i)],8);
                }
            }
          else
            { m &= ~dx_eq_mask16(v[j],(uint32_t) rc);
              const int32_t mylast = m ? p0 + (31 - __clz(m)) : -1;
              const int32_t inc = dx_warp_incl_max(mylast,lane);
              int32_t before = __shfl_up_sync(DX_FULL,inc,1);
              if (lane == 0) before = -1;
              pv = max(prev,before);
              prev = max(prev,__shfl_sync(DX_FULL,inc,31));
              uint32_t mm = m;
              int32_t  q = pv;
              while (mm)
                { const int i = __ffs(mm) - 1; mm &= mm - 1;
                  const int32_t p = p0 + i, r = p - q - 1;
                  lane_bits += item_len(run[min(r,255)],16) + item_len(sym[dx_byte_of(v[j],i)],8);
                  q = p;
                }
            }

          const uint32_t inc_bits = dx_warp_incl_sum(lane_bits,lane);
          const uint32_t row_bits = __shfl_sync(DX_FULL,inc_bits,31);
          if (MODE == 1 && stage_bits + row_bits > (uint32_t) (kStageWords-4)*32u)
            { // make room: flush the whole words staged so far
              const uint32_t nfull = stage_bits >> 5;
              __syncwarp();
              dx_warp_copy_out(gptr,stage,nfull*4u,lane);
              gptr += nfull*4u;
              __syncwarp();
              const uint32_t partial = stage[nfull];
              __syncwarp();
              for (uint32_t i = lane; i <= nfull; i += 32) stage[i] = (i == 0) ? partial : 0u;
              __syncwarp();
              flushed_bits += nfull*32u;
              stage_bits   &= 31u;
            }
          uint32_t pos = stage_bits + inc_bits - lane_bits;         // lane's first bit in stage

          // second walk over the lane's symbols: positions (+ emission in MODE 1)
          BitSink sink;
          if (MODE == 1) sink.start(stage,pos);
          if (rc < 0)
            { uint32_t mm = m;
              while (mm)
                { const int i = __ffs(mm) - 1; mm &= mm - 1;
                  const uint32_t x = dx_byte_of(v[j],i), e = sym[x];
                  const uint32_t len = DX_ENC_LEN(e);
                  if (MODE == 1) sink.put(DX_ENC_CODE(e),len);
                  int32_t here = (int32_t) (flushed_bits + pos);
                  pos += len;
                  if (DX_ENC_ESC(e))
                    { if (MODE == 1) sink.put(x,8);
                      here = (int32_t) (flushed_bits + pos);
                      pos += 8;
                    }
                  if (p0 + i == rlen-1) plast = here;
                }
            }
          else
            { uint32_t mm = m;
              int32_t  q = pv;
              while (mm)
                { const int i = __ffs(mm) - 1; mm &= mm - 1;
                  const int32_t p = p0 + i, r = p - q - 1;
                  q = p;
                  uint32_t e = run[min(r,255)], len = DX_ENC_LEN(e);
                  if (MODE == 1) sink.put(DX_ENC_CODE(e),len);
                  pos += len;
                  if (DX_ENC_ESC(e))
                    { if (MODE == 1) sink.put((uint32_t) r & 0xffffu,16);
                      pos += 16;
                    }
                  const uint32_t x = dx_byte_of(v[j],i);
                  e = sym[x]; len = DX_ENC_LEN(e);
                  if (MODE == 1) sink.put(DX_ENC_CODE(e),len);
                  int32_t here = (int32_t) (flushed_bits + pos);
                  pos += len;
                  if (DX_ENC_ESC(e))
                    { if (MODE == 1) sink.put(x,8);
                      here = (int32_t) (flushed_bits + pos);
                      pos += 8;
                    }
                  if (p == rlen-1) plast = here;
                }
            }
          if (MODE == 1) sink.done();
          stage_bits += row_bits;
        }
    }

  // trailing run of the run character (QV.c:475-487 when k reaches rlen inside a run)
  if (rc >= 0 && prev < rlen-1)
    { const int32_t r = rlen-1-prev;
      const uint32_t e = run[min(r,255)], len = DX_ENC_LEN(e);
      if (MODE == 1)
        { __syncwarp();
          if (lane == 0)
            { dx_or_bits(stage,stage_bits,DX_ENC_CODE(e),len);
              if (DX_ENC_ESC(e)) dx_or_bits(stage,stage_bits+len,(uint32_t) r & 0xffffu,16);
            }
        }
      plast = (int32_t) (flushed_bits + stage_bits + (DX_ENC_ESC(e) ? len : 0u));
      stage_bits += item_len(e,16);
    }

  // last item position: exactly one lane (or all, for the trailing run) holds it
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    plast = max(plast,__shfl_xor_sync(DX_FULL,plast,d));

  total_bits = flushed_bits + stage_bits;
  plast_out  = (uint32_t) plast;

  if (MODE == 1)
    { // final flush incl. the look-ahead padding word (QV.c:436-442)
      __syncwarp();
      uint32_t nst = (stage_bits + 31u) >> 5;
      const uint32_t full_total = (total_bits + 31u) >> 5;
      const uint32_t want_total = (rlen > 0) ? (((uint32_t) plast + 47u) >> 5) : 0u;
      if (want_total > full_total)
        { if (lane == 0)
            stage[nst] = (total_bits & 31u) ? stage[nst-1] : 0u;
          nst += 1;
          __syncwarp();
        }
      dx_warp_copy_out(gptr,stage,nst*4u,lane);
      gptr += nst*4u;
      __syncwarp();
      for (uint32_t i = lane; i <= nst; i += 32) stage[i] = 0u;
      __syncwarp();
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
  uint32_t kept_total = 0;       // symbols kept so far (all rows)
  uint32_t stage_syms = 0;       // symbols staged (2 bits each, from word 0)
  for (int32_t c0 = 0; c0 < lw.nchunk; c0 += 32)
    { const int32_t c = c0 + lane;
      const int32_t p0 = c*16 - lw.skew;
      uint32_t m = lw.valid(c);
      uint4 tv = make_uint4(0,0,0,0);
      if (c < lw.nchunk)
        { tv = dx_ldg16(lw.base + (int64_t) c*16);
          if (a.delchar >= 0)
            { uint4 dv = dx_ld16_any(del + p0,a.text_end16);    // del bytes at the same positions
              m &= ~dx_eq_mask16(dv,(uint32_t) a.delchar);
            }
        }
      const uint32_t cnt = __popc(m);
      const uint32_t inc = dx_warp_incl_sum(cnt,lane);
      const uint32_t row = __shfl_sync(DX_FULL,inc,31);
      if (MODE == 1)
        { if ((stage_syms + row)*2u > (uint32_t) (kStageWords-4)*32u)
            { const uint32_t nfull = (stage_syms*2u) >> 5;
              __syncwarp();
              dx_warp_copy_out(gptr,stage,nfull*4u,lane);
              gptr += nfull*4u;
              __syncwarp();
              const uint32_t partial = stage[nfull];
              __syncwarp();
              for (uint32_t i = lane; i <= nfull; i += 32) stage[i] = (i == 0) ? partial : 0u;
              __syncwarp();
              stage_syms &= 15u;
            }
          if (cnt)
            { uint32_t val = 0;
              uint32_t mm = m;
              while (mm)
                { const int i = __ffs(mm) - 1; mm &= mm - 1;
                  val = (val << 2) | base2(dx_byte_of(tv,i));
                }
              // MSB-first inside BYTES: compose big-endian words, store them byte-swapped
              const uint32_t pos = (stage_syms + inc - cnt)*2u, len = cnt*2u;
              const uint32_t w = pos >> 5, off = pos & 31u;
              if (off + len <= 32u)
                atomicOr(&stage[w],__byte_perm(val << (32u - off - len),0,0x0123));
              else
                { const uint32_t spill = off + len - 32u;
                  atomicOr(&stage[w],__byte_perm(val >> spill,0,0x0123));
                  atomicOr(&stage[w+1],__byte_perm(val << (32u - spill),0,0x0123));
                }
            }
          stage_syms += row;
        }
      kept_total += row;
    }
  if (MODE == 1)
    { __syncwarp();
      const uint32_t nbytes = (stage_syms + 3u) >> 2;      // partial word: only the bytes in use
      dx_warp_copy_out(gptr,stage,nbytes,lane);
      __syncwarp();
      for (uint32_t i = lane; i <= (nbytes >> 2) + 1; i += 32) stage[i] = 0u;
      __syncwarp();
    }
  return kept_total;
}

__device__ __forceinline__ uint32_t well_bytes(const EncArgs &a, int64_t e)
{ const int32_t lw = (e == 0) ? a.lwell_in : a.ent.well[e-1];
  const int32_t d  = a.ent.well[e] - lw;
  return 1u + (d >= 255 ? (uint32_t) d / 255u : 0u);            // dexqv.c:128-135
}

template <int MODE>
__global__ void __launch_bounds__(kEncThreads)
k_qv_code(EncArgs a)
{ extern __shared__ uint32_t smem[];
  uint32_t *stab  = smem;                                          // [6][256]
  uint32_t *stage = smem + 6*256 + (threadIdx.x >> 5)*kStageWords; // per warp (MODE 1 only)
  for (int i = threadIdx.x; i < 6*256; i += kEncThreads) stab[i] = a.tab[i];
  if (MODE == 1)
    for (int i = threadIdx.x; i < kEncWarps*kStageWords; i += kEncThreads) smem[6*256 + i] = 0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t nunits = a.ent.n * 5;
  while (true)
    { unsigned long long u0 = 0;
      if (lane == 0) u0 = atomicAdd(a.ticket,(unsigned long long) kFetch);
      u0 = __shfl_sync(DX_FULL,u0,0);
      if ((int64_t) u0 >= nunits) break;
      for (int f = 0; f < kFetch; f++)
        { const int64_t u = (int64_t) u0 + f;
          if (u >= nunits) break;
          const int64_t e = u / 5;
          const int     s = (int) (u - e*5);                       // 0 del 1 tag 2 ins 3 mrg 4 sub
          const int32_t rlen = a.ent.rlen[e];
          const uint8_t *l0  = a.text + a.ent.line0[e];
          const uint8_t *line = l0 + (int64_t) s*((int64_t) rlen + 1);
          uint8_t *gptr = NULL;
          if (MODE == 1)
            { int64_t o = a.off[e];
              for (int k = 0; k <= s; k++) o += a.bytes[e*6 + k];
              gptr = a.out + o;
            }
          if (s == 1)
            { uint32_t kept = code_tags<MODE>(a,l0,line,rlen,lane,stage,gptr);
              if (MODE == 0 && lane == 0) a.bytes[e*6 + 2] = (kept + 3u) >> 2;
            }
          else
            { uint32_t bits, plast;
              code_stream<MODE>(a,stab,line,rlen,s,lane,stage,gptr,bits,plast);
              if (MODE == 0 && lane == 0)
                { a.bytes[e*6 + 1 + s] = stream_words(bits,plast,rlen)*4u;
                  if (s == 0) a.bytes[e*6] = well_bytes(a,e) + 12u;
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
        }
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
  DX_CUDA(ctx,cudaMemcpyAsync(d_tab,h_tab,sizeof(QvEncTables),cudaMemcpyHostToDevice,ctx->stream));
  DX_CUDA(ctx,cudaMemsetAsync(d_ticket,0,16,ctx->stream));

  EncArgs a;
  a.text = d_text;
  a.text_end16 = d_text + ((text_n + 15) & ~(size_t) 15);
  a.ent = ent; a.tab = d_tab;
  a.delchar = delchar; a.subchar = subchar; a.lossy = lossy; a.lwell_in = lwell_in;
  a.bytes = d_bytes; a.off = d_off; a.out = d_out; a.ticket = d_ticket;

  const int grid = ctx->sm_count * 3;
  const size_t smem0 = 6*256*4;
  const size_t smem1 = 6*256*4 + (size_t) kEncWarps*kStageWords*4;
  DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_code<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem1));

  DX_PROF_BEGIN(ctx); k_qv_code<0><<<grid,kEncThreads,smem0,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,"k_qv_size");
  DX_PROF_BEGIN(ctx); k_qv_offsets<<<1,1024,0,ctx->stream>>>(d_bytes,n,d_off);
  DX_LAUNCHED(ctx,"k_qv_offsets");

  int64_t total = 0;
  int32_t lastw = 0;
  DX_CUDA(ctx,cudaMemcpyAsync(&total,d_off+n,8,cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaMemcpyAsync(&lastw,ent.well+(n-1),4,cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  if ((size_t) total > cap)
    return dx_fail(ctx,DX_E_CAP,"output needs %lld bytes, buffer has %zu",(long long) total,cap);

  a.ticket = d_ticket + 1;
  DX_PROF_BEGIN(ctx); k_qv_code<1><<<grid,kEncThreads,smem1,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,"k_qv_emit");
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
