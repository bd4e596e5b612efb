// dx_qv_stats.cu -- pass 1 of the QV coder on the device.
//
// Replaces Histogram_Seqs / Histogram_Runs (reference QV.c:702-724) and the order-dependent
// part of QVcoding_Scan (QV.c:988-1017):
//   k_find_delchar  first 'n'/'N' tag in file order -> delChar and the first entry whose
//                   deletion runs are counted (QV.c:993-1004)
//   k_probe_sub     walks entries until 100000 positions have been seen, then fixes subChar as
//                   the arg-max of the substitution histogram SO FAR (QV.c:1005-1015)
//   k_qv_hist       the four symbol histograms and the two run-length histograms, one warp per
//                   (entry, stream) line, thread-private byte counters in shared memory (plain
//                   read-modify-write, no atomics), flushed per stream with 64-bit global atomics

#include "dx_internal.h"
#include "dx_common.cuh"

namespace {

constexpr int kFetch       = 4;          // entries (lines of one stream) claimed per ticket

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
//
// Shared-memory atomics retire about two lanes per clock per SM, which made the first version of
// this kernel ATOMS-bound.  Here every THREAD owns a private set of one-byte counters in shared
// memory and increments them with plain load / add / store (no atomics, and no bank conflicts:
// counter b of thread t lives in word (b>>2)*T + t, byte b&3).  A counter that wraps to 0 adds
// 256 to a CTA-wide 64-bit histogram (one rare atomic per 256 hits); what is left in the byte
// counters is summed when the CTA moves on to the next stream and at the end.
//
//   RUN = false  streams without a run character: every byte is counted.  768 threads x 256 B.
//   RUN = true   del / sub with a run character: run bytes are skipped by a SWAR compare (their
//                count is recovered by subtraction on the host) and, from entry e_del / e_sub on,
//                the run length before every other symbol is counted too.  384 threads x 512 B.
//
// Work is handed out as (stream, block of kFetch entries) tickets in stream-major order, one line
// per warp at a time; a warp whose ticket belongs to a later stream waits at a CTA barrier until
// all warps have left the current stream, then the CTA flushes its counters.

struct HistArgs
{ const uint8_t *text;
  QvEntries      ent;
  int32_t        ns;               // streams handled by this launch
  int32_t        sidx[4];          // 0 del, 1 ins, 2 mrg, 3 sub
  int32_t        rc[4];            // run character of the stream (RUN launches)
  int64_t        efirst[4];        // first entry whose runs are counted
  unsigned long long *ticket;
  unsigned long long *ghist;       // [6][256]
};

template <bool RUN> struct HistCfg
{ static constexpr int kThreads = RUN ? 384 : 768;
  static constexpr int kBins    = RUN ? 512 : 256;
  static constexpr int kRows    = kBins / 4;
  static constexpr size_t kSmem = (size_t) kRows * kThreads * 4 + (size_t) kBins * 8;
};

template <int T>
__device__ __forceinline__ void bump(uint8_t *mine, unsigned long long *wide, uint32_t b)
{ uint8_t *p = mine + (b & ~3u) * T + (b & 3u);
  const uint32_t c = (uint32_t) *p + 1u;
  *p = (uint8_t) c;
  if (c == 256u) atomicAdd(&wide[b],256ull);
}

template <bool RUN>
__device__ void hist_flush(const HistArgs &a, int s, uint32_t *cnt, unsigned long long *wide)
{ typedef HistCfg<RUN> C;
  constexpr int G = C::kThreads / C::kRows;            // threads per row of 4 bins
  __syncthreads();
  { const int r = threadIdx.x / G, k0 = threadIdx.x % G;
    uint32_t a0 = 0, a1 = 0;
    uint32_t *row = cnt + (size_t) r * C::kThreads;
    for (int k = k0; k < C::kThreads; k += G)
      { const uint32_t w = row[k];
        row[k] = 0;
        a0 += w & 0x00ff00ffu;
        a1 += (w >> 8) & 0x00ff00ffu;
      }
    if (a0 & 0xffffu) atomicAdd(&wide[4*r],  (unsigned long long) (a0 & 0xffffu));
    if (a1 & 0xffffu) atomicAdd(&wide[4*r+1],(unsigned long long) (a1 & 0xffffu));
    if (a0 >> 16)     atomicAdd(&wide[4*r+2],(unsigned long long) (a0 >> 16));
    if (a1 >> 16)     atomicAdd(&wide[4*r+3],(unsigned long long) (a1 >> 16));
  }
  __syncthreads();
  for (int b = threadIdx.x; b < C::kBins; b += C::kThreads)
    { const unsigned long long v = wide[b];
      wide[b] = 0;
      if (v)
        { const int tab = (b < 256) ? s : (s == 0 ? 4 : 5);
          atomicAdd(&a.ghist[tab*256 + (b & 255)],v);
        }
    }
  __syncthreads();
}

template <bool RUN>
__global__ void __launch_bounds__(HistCfg<RUN>::kThreads,1)
k_qv_hist(HistArgs a)
{ typedef HistCfg<RUN> C;
  constexpr int T = C::kThreads;
  extern __shared__ __align__(16) uint8_t dx_hist_smem[];
  uint32_t *cnt = reinterpret_cast<uint32_t *>(dx_hist_smem);                  // [kRows][T] words
  unsigned long long *wide = reinterpret_cast<unsigned long long *>(dx_hist_smem + (size_t) C::kRows*T*4);
  for (int i = threadIdx.x; i < C::kRows*T; i += T) cnt[i] = 0;
  for (int i = threadIdx.x; i < C::kBins; i += T) wide[i] = 0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  uint8_t *mine = dx_hist_smem + 4*threadIdx.x;
  const int lineidx[4] = { 0, 2, 3, 4 };
  const int64_t nb = (a.ent.n + kFetch - 1) / kFetch;
  const int64_t total = nb * a.ns;
  int cur = 0;

  while (true)
    { unsigned long long u = 0;
      if (lane == 0) u = atomicAdd(a.ticket,1ull);
      u = __shfl_sync(DX_FULL,u,0);
      const int si = ((int64_t) u >= total) ? a.ns : (int) ((int64_t) u / nb);
      while (cur < si)                                   // every warp passes every stream boundary
        { hist_flush<RUN>(a,a.sidx[cur],cnt,wide);
          cur++;
        }
      if (si >= a.ns) break;
      const int s = a.sidx[si];
      const int64_t e0 = ((int64_t) u - (int64_t) si*nb) * kFetch;
      const int64_t e1 = min(a.ent.n,e0 + kFetch);
      for (int64_t e = e0; e < e1; e++)
        { const int32_t rlen = a.ent.rlen[e];
          if (rlen == 0) continue;
          const uint8_t *line = a.text + a.ent.line0[e] + (int64_t) lineidx[s]*((int64_t) rlen + 1);
          const int skew = (int) (reinterpret_cast<uintptr_t>(line) & 15);
          const uint8_t *base = line - skew;                       // 16-byte aligned
          const int32_t nchunk = (skew + rlen + 15) >> 4;
          const uint32_t rc = RUN ? (uint32_t) a.rc[si] : 0u;
          const bool runs = RUN && (e >= a.efirst[si]);
          int32_t prev = -1;                                       // last non-run position so far

          for (int32_t c0 = 0; c0 < nchunk; c0 += 128)
            { uint4 v[4];
#pragma unroll
              for (int j = 0; j < 4; j++)
                { const int32_t c = c0 + j*32 + lane;
                  v[j] = (c < nchunk) ? dx_ldg16(base + (int64_t) c*16) : make_uint4(0,0,0,0);
                }
#pragma unroll
              for (int j = 0; j < 4; j++)
                { const int32_t c = c0 + j*32 + lane;
                  if (c0 + j*32 >= nchunk) break;                    // warp-uniform
                  const int32_t p0 = c*16 - skew;                   // line position of byte 0
                  uint32_t valid = 0;
                  if (c < nchunk)
                    valid = dx_range16(max(0,-p0),min(16,rlen - p0));
                  if (!RUN)
                    { const uint32_t w[4] = { v[j].x, v[j].y, v[j].z, v[j].w };
                      if (valid == 0xffffu)
                        {
#pragma unroll
                          for (int i = 0; i < 16; i++)
                            bump<T>(mine,wide,(w[i >> 2] >> ((i & 3)*8)) & 0xffu);
                        }
                      else
                        { uint32_t m = valid;
                          while (m)
                            { const int i = __ffs(m) - 1; m &= m - 1;
                              bump<T>(mine,wide,dx_byte_of(v[j],i));
                            }
                        }
                    }
                  else
                    { uint32_t m = valid & ~dx_eq_mask16(v[j],rc);         // bytes that are not the run character
                      int32_t pv = -1;
                      if (runs)
                        { const int32_t mylast = m ? p0 + (31 - __clz(m)) : -1;
                          const int32_t inc = dx_warp_incl_max(mylast,lane);
                          int32_t before = __shfl_up_sync(DX_FULL,inc,1);
                          if (lane == 0) before = -1;
                          pv = max(prev,before);
                          prev = max(prev,__shfl_sync(DX_FULL,inc,31));
                        }
                      while (m)
                        { const int i = __ffs(m) - 1; m &= m - 1;
                          bump<T>(mine,wide,dx_byte_of(v[j],i));
                          if (runs)
                            { const int32_t p = p0 + i;
                              bump<T>(mine,wide,256u + (uint32_t) min(p - pv - 1,255));
                              pv = p;
                            }
                        }
                    }
                }
            }
          if (runs && lane == 0 && prev < rlen-1)                  // trailing run (QV.c:713-720)
            bump<T>(mine,wide,256u + (uint32_t) min(rlen-1-prev,255));
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
  unsigned long long *d_hist = (unsigned long long *) dx_arena_get(ctx,6*256*8 + 16);
  if (d_hist == NULL) return DX_E_NOMEM;
  DX_CUDA(ctx,cudaMemsetAsync(d_hist,0,6*256*8 + 16,ctx->stream));
  HistArgs plain, run;
  memset(&plain,0,sizeof(plain)); memset(&run,0,sizeof(run));
  plain.text = run.text = d_text; plain.ent = run.ent = ent;
  plain.ghist = run.ghist = d_hist;
  plain.ticket = d_hist + 6*256; run.ticket = d_hist + 6*256 + 1;
  for (int s = 0; s < 4; s++)
    { const int32_t rc = (s == 0) ? h_probe->delchar : (s == 3) ? h_probe->subchar : -1;
      HistArgs &h = (rc >= 0) ? run : plain;
      h.sidx[h.ns] = s; h.rc[h.ns] = rc;
      h.efirst[h.ns] = (s == 0) ? h_probe->e_del : h_probe->e_sub;
      h.ns++;
    }
  static bool attr_done = false;
  if (!attr_done)
    { DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_hist<false>,cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) HistCfg<false>::kSmem));
      DX_CUDA(ctx,cudaFuncSetAttribute(k_qv_hist<true>,cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) HistCfg<true>::kSmem));
      attr_done = true;
    }
  if (plain.ns > 0)
    { DX_PROF_BEGIN(ctx);
      k_qv_hist<false><<<ctx->sm_count,HistCfg<false>::kThreads,HistCfg<false>::kSmem,ctx->stream>>>(plain);
      DX_LAUNCHED(ctx,"k_qv_hist_plain");
    }
  if (run.ns > 0)
    { DX_PROF_BEGIN(ctx);
      k_qv_hist<true><<<ctx->sm_count,HistCfg<true>::kThreads,HistCfg<true>::kSmem,ctx->stream>>>(run);
      DX_LAUNCHED(ctx,"k_qv_hist_run");
    }
  DX_CUDA(ctx,cudaMemcpyAsync(h_hist,d_hist,6*256*8,cudaMemcpyDeviceToHost,ctx->stream));
  DX_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  return DX_OK;
}
