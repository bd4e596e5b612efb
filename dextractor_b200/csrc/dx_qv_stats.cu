// dx_qv_stats.cu -- pass 1 of the QV coder on the device.
//
// Replaces Histogram_Seqs / Histogram_Runs (reference QV.c:702-724) and the order-dependent
// part of QVcoding_Scan (QV.c:988-1017):
//   k_find_delchar  first 'n'/'N' tag in file order -> delChar and the first entry whose
//                   deletion runs are counted (QV.c:993-1004)
//   k_probe_sub     walks entries until 100000 positions have been seen, then fixes subChar as
//                   the arg-max of the substitution histogram SO FAR (QV.c:1005-1015)
//   k_qv_hist       the four symbol histograms and the two run-length histograms, one warp per
//                   (entry, stream) line, lane-replicated shared-memory bins (no intra-warp
//                   bank or address conflicts), flushed once per CTA with 64-bit global atomics

#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kHistThreads = 1024;
constexpr int kRunRep      = 8;          // replicas of the run-length bins
constexpr int kFetch       = 8;          // lines claimed per atomic ticket

// ---- delChar ---------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
k_find_delchar(const uint8_t *text, QvEntries ent, unsigned long long *found)
{ const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t) gridDim.x * blockDim.x) >> 5;
  for (int64_t e = warp; e < ent.n; e += nwarp)
    { if ((*(volatile unsigned long long *) found >> 32) < (unsigned long long) e)
        return;                                    // an earlier entry already has one
      const int32_t rlen = ent.rlen[e];
      const uint8_t *tag = text + ent.line0[e] + (int64_t) rlen + 1;
      for (int32_t b = 0; b < rlen; b += 32)
        { int32_t k = b + lane;
          bool hit = (k < rlen) && (tag[k] == 'n' || tag[k] == 'N');
          uint32_t m = __ballot_sync(DX_FULL,hit);
          if (m)
            { if (lane == 0)
                atomicMin(found,((unsigned long long) e << 32) | (uint32_t) (b + __ffs(m) - 1));
              return;
            }
        }
    }
}

__global__ void k_pick_delchar(const uint8_t *text, QvEntries ent, const unsigned long long *found,
                               QvProbe *probe)
{ unsigned long long f = *found;
  if (f == ~0ull)
    { probe->delchar = -1; probe->e_del = ent.n; }
  else
    { int64_t e = (int64_t) (f >> 32);
      probe->delchar = text[ent.line0[e] + (uint32_t) f];
      probe->e_del   = e;
    }
}

// ---- subChar ---------------------------------------------------------------------------------

__global__ void __launch_bounds__(1024)
k_probe_sub(const uint8_t *text, QvEntries ent, uint64_t tot_in, const uint64_t *sub_in,
            QvProbe *probe)
{ __shared__ uint32_t h[256];
  __shared__ unsigned long long key[256];
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
  uint64_t tot = tot_in;
  int64_t  e = 0;
  bool     fixed = false;
  for ( ; e < ent.n; e++)
    { const int32_t rlen = ent.rlen[e];
      const uint8_t *sub = text + ent.line0[e] + 4*((int64_t) rlen + 1);
      for (int32_t k = threadIdx.x; k < rlen; k += blockDim.x)
        atomicAdd(&h[sub[k]],1u);
      tot += (uint64_t) rlen;
      if (tot >= 100000) { fixed = true; break; }
    }
  __syncthreads();
  if (threadIdx.x < 256)
    { uint64_t c = sub_in[threadIdx.x] + h[threadIdx.x];
      probe->sub_prefix[threadIdx.x] = c;
      // arg-max with the FIRST maximum winning: order by (count desc, symbol asc)
      key[threadIdx.x] = (c << 8) | (255u - threadIdx.x);
    }
  __syncthreads();
  if (threadIdx.x == 0)
    { unsigned long long best = 0;
      for (int k = 0; k < 256; k++) best = max(best,key[k]);
      probe->subchar = fixed ? (int32_t) (255u - (uint32_t) (best & 0xff)) : -1;
      probe->e_sub   = fixed ? e : ent.n;
      probe->totchar = tot;
    }
}

// ---- histograms ------------------------------------------------------------------------------

struct HistArgs
{ const uint8_t *text;
  QvEntries      ent;
  int32_t        delchar, subchar;
  int64_t        e_del, e_sub;
  unsigned long long *ticket;
  unsigned long long *ghist;       // [6][256]
};

// one 16-byte chunk of one line, bytes outside `valid` ignored
__device__ __forceinline__ void hist_chunk(uint4 v, uint32_t valid, int32_t rc, uint32_t *sh,
                                           int lane)
{ if (rc < 0 && valid == 0xffffu)
    { const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for (int i = 0; i < 16; i++)
        atomicAdd(&sh[((w[i >> 2] >> ((i & 3)*8)) & 0xffu)*32 + lane],1u);
      return;
    }
  uint32_t m = valid;
  if (rc >= 0) m &= ~dx_eq_mask16(v,(uint32_t) rc);
  while (m)
    { int i = __ffs(m) - 1;
      m &= m - 1;
      atomicAdd(&sh[dx_byte_of(v,i)*32 + lane],1u);
    }
}

__global__ void __launch_bounds__(kHistThreads,1)
k_qv_hist(HistArgs a)
{ extern __shared__ uint32_t smem[];
  uint32_t *sh = smem;                             // [4][256][32]
  uint32_t *rh = smem + 4*256*32;                  // [2][256][kRunRep]
  for (int i = threadIdx.x; i < 4*256*32 + 2*256*kRunRep; i += kHistThreads) smem[i] = 0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t nunits = a.ent.n * 4;
  const int lineidx[4] = { 0, 2, 3, 4 };

  while (true)
    { unsigned long long u0 = 0;
      if (lane == 0) u0 = atomicAdd(a.ticket,(unsigned long long) kFetch);
      u0 = __shfl_sync(DX_FULL,u0,0);
      if ((int64_t) u0 >= nunits) break;
      for (int f = 0; f < kFetch; f++)
        { const int64_t u = (int64_t) u0 + f;
          if (u >= nunits) break;
          const int64_t e = u >> 2;
          const int     s = (int) (u & 3);
          const int32_t rlen = a.ent.rlen[e];
          if (rlen == 0) continue;
          const int32_t rc = (s == 0) ? a.delchar : (s == 3) ? a.subchar : -1;
          const bool runs  = (s == 0 && a.delchar >= 0 && e >= a.e_del) ||
                             (s == 3 && a.subchar >= 0 && e >= a.e_sub);
          const uint8_t *line = a.text + a.ent.line0[e] + (int64_t) lineidx[s]*((int64_t) rlen + 1);
          const int skew = (int) (reinterpret_cast<uintptr_t>(line) & 15);
          const uint8_t *base = line - skew;                       // 16-byte aligned
          const int32_t nchunk = (skew + rlen + 15) >> 4;
          uint32_t *hs = sh + s*256*32;
          uint32_t *hr = rh + (s == 3 ? 256*kRunRep : 0);
          int32_t prev = -1;                                       // last non-run position so far

          for (int32_t c0 = 0; c0 < nchunk; c0 += 128)
            { uint4 v[4];
#pragma unroll
              for (int j = 0; j < 4; j++)
                { int32_t c = c0 + j*32 + lane;
                  v[j] = (c < nchunk) ? dx_ldg16(base + (int64_t) c*16) : make_uint4(0,0,0,0);
                }
#pragma unroll
              for (int j = 0; j < 4; j++)
                { const int32_t c = c0 + j*32 + lane;
                  const int32_t p0 = c*16 - skew;                   // line position of byte 0
                  uint32_t valid = 0;
                  if (c < nchunk)
                    valid = dx_range16(max(0,-p0),min(16,rlen - p0));
                  hist_chunk(v[j],valid,rc,hs,lane);
                  if (runs)
                    { uint32_t m = valid & ~dx_eq_mask16(v[j],(uint32_t) rc);   // non-run bytes
                      int32_t mylast = m ? p0 + (31 - __clz(m)) : -1;
                      int32_t inc = dx_warp_incl_max(mylast,lane);
                      int32_t before = __shfl_up_sync(DX_FULL,inc,1);
                      if (lane == 0) before = -1;
                      int32_t pv = max(prev,before);
                      while (m)
                        { int i = __ffs(m) - 1;
                          m &= m - 1;
                          int32_t p = p0 + i;
                          int32_t r = p - pv - 1;
                          atomicAdd(&hr[min(r,255)*kRunRep + (lane & (kRunRep-1))],1u);
                          pv = p;
                        }
                      prev = max(prev,__shfl_sync(DX_FULL,inc,31));
                    }
                }
            }
          if (runs && lane == 0 && prev < rlen-1)                  // trailing run (QV.c:713-720)
            atomicAdd(&hr[min(rlen-1-prev,255)*kRunRep],1u);
        }
    }
  __syncthreads();

  // flush: thread t owns (stream t>>8, bin t&255); rotate replica reads to dodge bank conflicts
  { const int s = threadIdx.x >> 8, bin = threadIdx.x & 255;
    uint32_t sum = 0;
#pragma unroll 8
    for (int r = 0; r < 32; r++)
      sum += sh[(s*256 + bin)*32 + ((r + lane) & 31)];
    if (sum) atomicAdd(&a.ghist[s*256 + bin],(unsigned long long) sum);
    if (threadIdx.x < 512)
      { const int q = threadIdx.x >> 8;
        uint32_t rs = 0;
        for (int r = 0; r < kRunRep; r++)
          rs += rh[(q*256 + bin)*kRunRep + r];
        if (rs) atomicAdd(&a.ghist[(4+q)*256 + bin],(unsigned long long) rs);
      }
  }
}

}  // namespace

int dxk_qv_probe(dx_ctx *ctx, const uint8_t *d_text, QvEntries ent, const dx_qv_carry *carry,
                 QvProbe *h_probe)
{ QvProbe *d_probe = (QvProbe *) dx_arena_get(ctx,sizeof(QvProbe));
  unsigned long long *d_found = (unsigned long long *) dx_arena_get(ctx,8);
  uint64_t *d_sub_in = (uint64_t *) dx_arena_get(ctx,256*8);
  if (d_probe == NULL || d_found == NULL || d_sub_in == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_probe,0,sizeof(QvProbe),ctx->stream));
  DX_CUDA(ctx,cudaMemsetAsync(d_found,0xff,8,ctx->stream));
  DX_CUDA(ctx,cudaMemcpyAsync(d_sub_in,carry->sub,256*8,cudaMemcpyHostToDevice,ctx->stream));

  const bool need_del = (carry->delchar < 0), need_sub = (carry->subchar < 0);
  if (need_del && ent.n > 0)
    { int blocks = ctx->sm_count * 4;
      DX_PROF_BEGIN(ctx); k_find_delchar<<<blocks,256,0,ctx->stream>>>(d_text,ent,d_found);
      DX_LAUNCHED(ctx,"k_find_delchar");
      DX_PROF_BEGIN(ctx); k_pick_delchar<<<1,1,0,ctx->stream>>>(d_text,ent,d_found,d_probe);
      DX_LAUNCHED(ctx,"k_pick_delchar");
    }
  if (need_sub)
    { k_probe_sub<<<1,1024,0,ctx->stream>>>(d_text,ent,carry->totchar,d_sub_in,d_probe);
      DX_LAUNCHED(ctx,"k_probe_sub");
    }
  DX_CUDA(ctx,cudaMemcpyAsync(h_probe,d_probe,sizeof(QvProbe),cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  if (!need_del || ent.n == 0)
    { h_probe->delchar = carry->delchar; h_probe->e_del = (carry->delchar >= 0) ? 0 : ent.n; }
  if (!need_sub)
    { h_probe->subchar = carry->subchar; h_probe->e_sub = 0;
      memcpy(h_probe->sub_prefix,carry->sub,sizeof(carry->sub));
    }
  return DX_OK;
}

int dxk_qv_hist(dx_ctx *ctx, const uint8_t *d_text, QvEntries ent, const QvProbe *h_probe,
                uint64_t *h_hist, int32_t *h_newline_inside)
{ if (h_newline_inside) *h_newline_inside = 0;
  memset(h_hist,0,6*256*8);
  if (ent.n == 0) return DX_OK;
  unsigned long long *d_hist = (unsigned long long *) dx_arena_get(ctx,6*256*8 + 8);
  if (d_hist == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_hist,0,6*256*8 + 8,ctx->stream));
  HistArgs a;
  a.text = d_text; a.ent = ent;
  a.delchar = h_probe->delchar; a.subchar = h_probe->subchar;
  a.e_del = h_probe->e_del; a.e_sub = h_probe->e_sub;
  a.ghist = d_hist; a.ticket = d_hist + 6*256;
  const size_t smem = (4*256*32 + 2*256*kRunRep) * sizeof(uint32_t);
  static bool attr_done = false;
  if (!attr_done)
    { DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_hist,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem));
      attr_done = true;
    }
  DX_PROF_BEGIN(ctx); k_qv_hist<<<ctx->sm_count,kHistThreads,smem,ctx->stream>>>(a);
  DX_LAUNCHED(ctx,"k_qv_hist");
  DX_CUDA(ctx,cudaMemcpyAsync(h_hist,d_hist,6*256*8,cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  return DX_OK;
}
