// dx_qv_plan.cu -- device-side planning of a .dexqv decode.
//
// undexqv.c:119-208 interleaves three things per entry: reading the header fields, formatting the
// header line, and decoding the streams.  For a whole file the first two are prefix sums: the well
// number is the running sum of the well-delta bytes (undexqv.c:127-133), the place of an entry in
// the text is the running sum of header-line lengths and 5*(rlen+1).  These kernels compute them
// for all entries at once so that the host only has to verify the chain of entries.

#include "dx_internal.h"
#include "dx_common.cuh"
#include "dx_chain.h"

namespace {

__device__ __forceinline__ int32_t ld_le32(const uint8_t *p)
{ return (int32_t) ((uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24)); }

__device__ __forceinline__ uint32_t ndig(int32_t v)          // characters of printf("%d")
{ uint32_t n = (v < 0);
  uint32_t u = (v < 0) ? (uint32_t) (-(int64_t) v) : (uint32_t) v;
  n += (u >= 1000000000u) ? 10u : (u >= 100000000u) ? 9u : (u >= 10000000u) ? 8u : (u >= 1000000u) ? 7u
     : (u >= 100000u) ? 6u : (u >= 10000u) ? 5u : (u >= 1000u) ? 4u : (u >= 100u) ? 3u : (u >= 10u) ? 2u : 1u;
  return n;
}

// entry starts known: skip the 0xff bytes of the well delta, read the fields
__global__ void k_qv_known_prep(const uint8_t *in, int64_t n, const int64_t *estart, int64_t count,
                                QvPlanArrays pa, int32_t *flag)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int64_t p = estart[i];
  if (p < 0 || p >= n) { atomicExch(flag,1); pa.rlen[i] = 0; pa.delta[i] = 0; pa.fs[i] = 0; return; }
  while (p < n && in[p] == 0xff) p++;
  if (p + 13 > n) { atomicExch(flag,1); pa.rlen[i] = 0; pa.delta[i] = 0; pa.fs[i] = 0; return; }
  pa.delta[i] = (uint32_t) (255*(p - estart[i])) + in[p];
  const uint8_t *f = in + p + 1;
  const int32_t beg = ld_le32(f), en = ld_le32(f+4), qv = ld_le32(f+8);
  pa.beg[i] = beg; pa.end[i] = en; pa.qv[i] = qv;
  const int64_t rl = (int64_t) en - beg;
  if (rl < 0 || rl >= (1 << 24)) { atomicExch(flag,2); pa.rlen[i] = 0; }
  else pa.rlen[i] = (int32_t) rl;
  pa.fs[i] = p + 13;
}

// candidates: fields, context for the chain (0xff run before the terminator byte, the terminator),
// the byte limit of the speculative decode, the length of the lines in the scratch image
__global__ void k_qv_cand_prep(const uint8_t *in, int64_t n, int64_t first, const int64_t *q, int64_t count,
                               int span, int minbits, QvPlanArrays pa, uint32_t *tlen, int64_t *limit,
                               int32_t *ffrun, uint8_t *last)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t qi = q[i];
  const uint8_t *f = in + qi;
  const int32_t beg = ld_le32(f), en = ld_le32(f+4), qv = ld_le32(f+8);
  pa.beg[i] = beg; pa.end[i] = en; pa.qv[i] = qv;
  pa.fs[i] = qi + 12;
  int64_t rl = (int64_t) en - beg;
  if (rl < 0 || rl >= (1 << 24)) rl = 0;
  tlen[i] = (uint32_t) (5*(rl + 1));
  // context
  const int64_t p = qi - 1;
  if (p < first) { ffrun[i] = -1; last[i] = 0; pa.delta[i] = 0; }
  else
    { last[i] = in[p];
      int32_t r = 0;
      int64_t k = p - 1;
      while (k >= first && in[k] == 0xff && r < (1 << 20)) { r++; k--; }
      ffrun[i] = r;
      // the well delta IF this is an entry less than 255 wells after its predecessor (0xff bytes in
      // front of the terminator then belong to the previous stream, whose last byte is the top of a
      // code word and 0xff for one entry in 200): the usual case, in which the text can be decoded
      // straight into place, see undexqv_fast
      pa.delta[i] = in[p];
      // (the 0xff bytes in front of the FIRST entry of the image have no stream to belong to: a shard
      //  of a larger file starts with the delta against its predecessor's last well, which can be large)
      if (r > 0 && p - r == first) pa.delta[i] += 255u * (uint32_t) r;
    }
  // limit: the fields of the span-th candidate after this one, counting only candidates at least
  // 64 bytes after the previously counted one
  int64_t k = i, at = qi;
  for (int h = 0; h < span && k < count; h++)
    { int64_t j = k + 1;
      while (j < count && q[j] - at < 64) j++;
      k = j;
      if (k < count) at = q[k];
    }
  const int64_t lim = (k < count) ? q[k] : n;
  limit[i] = lim;
  if (rl > 0 && qi + 12 + ((rl*minbits) >> 3) > lim) pa.rlen[i] = -1;       // cannot fit: not decoded
  else pa.rlen[i] = (int32_t) rl;
}

// bytes of entry m in the text: header line + five lines.  cand (optional) maps entries to rows of
// the field arrays; well[m] = well_in + wpre[m+1] when wells come from a scan of the deltas
__global__ void k_qv_text_len(int64_t count, const int32_t *cand, QvPlanArrays pa, const int64_t *wpre,
                              const int32_t *wells, int32_t well_in, int plen, uint32_t *len,
                              int32_t *well_out, int32_t *flag, const int32_t *sel)
{ const int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= count) return;
  const int64_t c = cand ? cand[m] : m;
  const int32_t well = wells ? wells[m] : (int32_t) (well_in + wpre[m+1]);
  if (sel != NULL && sel[c] < 0) { len[m] = 0; well_out[m] = well; return; }     // not part of the layout
  const int32_t beg = pa.beg[c], en = pa.end[c], qv = pa.qv[c];
  const int64_t rl = (int64_t) en - beg;
  if (rl < 0 || rl >= (1 << 24)) { atomicExch(flag,2); len[m] = 0; well_out[m] = well; return; }
  const uint32_t hl = (uint32_t) plen + 1u + ndig(well) + 1u + ndig(beg) + 1u + ndig(en) + 6u + ndig(qv) + 1u;
  len[m] = hl + (uint32_t) (5*(rl + 1));
  well_out[m] = well;
}

// The layout of the text IF the entries are the candidates kept here (undexqv_fast).  Dropped:
//   * of candidates whose fields overlap (closer than 13 bytes: the second subread of a well, delta
//     0 behind zero padding, has look-alikes 4 and 8 bytes in front of it whose fields are the true
//     ones shifted by one or two words) all but the last -- two entries can never be that close;
//   * candidates outside what real headers hold (region score 0..1000, bax.c:347; a read starting
//     within the first 16 M bases of its well): the index's filter is wide on purpose (beg < 2^27,
//     qv < 2^16 -- about one hit per 2 GB file in the middle of stream data), this one is not.
// A dropped candidate leaves the layout: no text, no well delta, not decoded.  A true entry dropped
// here only sends the call down the scratch-image form.
__global__ void k_qv_direct_prep(const int64_t *q, int64_t count, QvPlanArrays pa, int32_t *rlen_d, uint8_t *keep)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const bool kept = !(i + 1 < count && q[i+1] - q[i] < 13) &&
                    (uint32_t) pa.qv[i] <= 1000u && (uint32_t) pa.beg[i] < (1u << 24);
  keep[i] = kept ? 1 : 0;
  rlen_d[i] = kept ? pa.rlen[i] : -1;
  if (!kept) pa.delta[i] = 0;
}

// ... and the check that the decode confirmed that layout (the host's chain walk, undexqv_fast, as a
// predicate over independent candidates): the first kept candidate's terminator byte is the first
// byte after the coding header (or only 0xff bytes lie between), every kept candidate decoded cleanly, has a terminator below 0xff
// and ends exactly where the terminator of the next kept candidate stands (no 0xff delta bytes),
// the last one ends at the end of the image.  flag[0] = 1 on any violation; flag[1] += kept.
__global__ void k_qv_chain_check(const int64_t *q, int64_t count, const uint8_t *keep, const int32_t *rlen_d,
                                 const int32_t *stat, const int64_t *soff, const uint8_t *last, const int32_t *ffrun,
                                 int64_t first, int64_t n, int32_t *flag)
{ const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || !keep[i]) return;
  atomicAdd(flag + 1,1);
  const DxChainIn c = { q, ffrun, last, stat, soff, keep, count, first, n };
  if (!dx_chain_check_one(c,rlen_d,i)) atomicExch(flag,1);
}

__global__ void k_qv_build_ent(int64_t count, const int32_t *cand, QvPlanArrays pa, const int32_t *well,
                               const int64_t *opre, const uint32_t *len, const int64_t *toff,
                               QvDecEntry *ent, int64_t *src, int64_t *fs_out, int32_t *rlen_out)
{ const int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= count) return;
  const int64_t c = cand ? cand[m] : m;
  QvDecEntry d;
  d.well = well[m]; d.beg = pa.beg[c]; d.end = pa.end[c]; d.qv = pa.qv[c];
  const int64_t rl = (int64_t) d.end - d.beg;
  d.out_off = opre[m];
  d.text_off = opre[m] + (int64_t) len[m] - 5*(rl + 1);
  ent[m] = d;
  if (src != NULL) src[m] = toff[c];
  if (fs_out != NULL) { fs_out[m] = pa.fs[c]; rlen_out[m] = (int32_t) rl; }
}

}  // namespace

int dxk_qv_known_prep(dx_ctx *ctx, const uint8_t *d_in, size_t n, const int64_t *d_estart, int64_t count,
                      QvPlanArrays pa, int32_t *d_flag)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx); k_qv_known_prep<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_in,(int64_t) n,d_estart,count,pa,d_flag);
  DX_LAUNCHED(ctx,"k_qv_known_prep");
  return DX_OK;
}

int dxk_qv_cand_prep(dx_ctx *ctx, const uint8_t *d_in, size_t n, size_t first, const int64_t *d_q, int64_t count,
                     int span, int minbits, QvPlanArrays pa, uint32_t *d_tlen, int64_t *d_limit,
                     int32_t *d_ffrun, uint8_t *d_last)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx);
  k_qv_cand_prep<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_in,(int64_t) n,(int64_t) first,d_q,count,span,
                                                                    minbits,pa,d_tlen,d_limit,d_ffrun,d_last);
  DX_LAUNCHED(ctx,"k_qv_cand_prep");
  return DX_OK;
}

int dxk_qv_text_len(dx_ctx *ctx, int64_t count, const int32_t *d_cand, QvPlanArrays pa, const int64_t *d_wpre,
                    const int32_t *d_wells, int32_t well_in, int plen, uint32_t *d_len, int32_t *d_well_out,
                    int32_t *d_flag, const int32_t *d_sel)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx);
  k_qv_text_len<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(count,d_cand,pa,d_wpre,d_wells,well_in,plen,d_len,
                                                                   d_well_out,d_flag,d_sel);
  DX_LAUNCHED(ctx,"k_qv_text_len");
  return DX_OK;
}

int dxk_qv_direct_prep(dx_ctx *ctx, const int64_t *d_q, int64_t count, QvPlanArrays pa, int32_t *d_rlen_d,
                       uint8_t *d_keep)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx);
  k_qv_direct_prep<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_q,count,pa,d_rlen_d,d_keep);
  DX_LAUNCHED(ctx,"k_qv_direct_prep");
  return DX_OK;
}

int dxk_qv_build_ent(dx_ctx *ctx, int64_t count, const int32_t *d_cand, QvPlanArrays pa, const int32_t *d_well,
                     const int64_t *d_opre, const uint32_t *d_len, const int64_t *d_toff, QvDecEntry *d_ent,
                     int64_t *d_src, int64_t *d_fs_out, int32_t *d_rlen_out)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx);
  k_qv_build_ent<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(count,d_cand,pa,d_well,d_opre,d_len,d_toff,d_ent,
                                                                    d_src,d_fs_out,d_rlen_out);
  DX_LAUNCHED(ctx,"k_qv_build_ent");
  return DX_OK;
}

int dxk_qv_chain_check(dx_ctx *ctx, const int64_t *d_q, int64_t count, const uint8_t *d_keep, const int32_t *d_rlen_d,
                       const int32_t *d_stat, const int64_t *d_soff, const uint8_t *d_last, const int32_t *d_ffrun,
                       size_t first, size_t n, int32_t *d_flag)
{ if (count == 0) return DX_OK;
  DX_PROF_BEGIN(ctx);
  k_qv_chain_check<<<(unsigned) ((count+255)/256),256,0,ctx->stream>>>(d_q,count,d_keep,d_rlen_d,d_stat,d_soff,d_last,d_ffrun,
                                                                      (int64_t) first,(int64_t) n,d_flag);
  DX_LAUNCHED(ctx,"k_qv_chain_check");
  return DX_OK;
}
