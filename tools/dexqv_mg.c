/* dexqv_mg -- dexqv of ONE .quiva file on N GPUs of one node: .quiva -> .dexqv, the same bytes the
 * reference's dexqv writes (dexqv.c:59-147), the file cut into N shards of whole entries.
 *
 *     dexqv_mg [-vklc] -g<N> <path:quiva> ...
 *
 * One process, one host thread per GPU, ncclCommInitAll.  What crosses NVLink:
 *   1. ncclAllGather   the newline count of every rank's byte range (a line that starts with '@'
 *                      proves nothing -- '@' is QV 31 -- so entry starts inside the file are found by
 *                      counting lines: entries are 6 lines, QV.c:751-798), then the entry-aligned starts;
 *   2. ncclBroadcast   the run characters the file's first ~100 000 positions fix (QV.c:993-1015); as
 *                      long as they are not fixed the shards are scanned in file order, each rank
 *                      handing its carry to the next, afterwards all remaining ranks scan at once;
 *   3. ncclAllReduce   the six 256-bin histograms + position and entry counts (uint64 sum); every rank
 *                      then builds the SAME code tables on its host (Create_QVcoding is deterministic);
 *   4. ncclAllGather   last well and compressed size of every shard: the first well delta of shard r
 *                      is coded against the last well of shard r-1 (dexqv.c:128-135) and its bytes
 *                      go to header + sum of the sizes before it (pwrite; the reference's implicit
 *                      file position).
 * No bulk data crosses NVLink.  -c: every rank decodes its shard again and compares it with the file
 * (meaningful for files whose deletion tags are 'n' exactly where the deletion QV is the run
 * character, as the extractor writes them: other tags do not survive the reference's coding either).
 * Host code is plain C; every codec step runs on the GPUs through libdexb200.so.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>
#include <nccl.h>
#include "dexb200.h"

#define MAXGPU   16
#define STAGE    ((size_t) 64 << 20)         /* pinned staging buffers (two per rank) */
#define SLACK    ((size_t) 96 << 20)         /* an entry is at most 5*(2^24+1) + header bytes */

static const char *Prog = "dexqv_mg";

typedef struct
  { int verbose, keep, lossy, check, world; } Opts;

typedef struct
  { int         rank;
    const Opts *o;
    ncclComm_t  comm;
    int         fd_in, fd_out;
    int64_t     fsize;
    double      t_read, t_scan, t_code, t_write;   /* seconds (rank 0 prints the maxima) */
    int64_t     nbytes, nent, obytes;
  } Rank;

static double now(void)
{ struct timeval tv; gettimeofday(&tv,NULL); return tv.tv_sec + 1e-6*tv.tv_usec; }

static void fail(dx_ctx *ctx, int rank, const char *what, int rc)
{ fprintf(stderr,"%s: rank %d: %s: %s (%d)\n",Prog,rank,what,ctx ? dx_strerror(ctx) : "",rc);
  exit (1);                                    /* one rank down: the whole tool stops (no partial file) */
}
#define DXC(call) do { int rc__ = (call); if (rc__ != DX_OK) fail(ctx,R->rank,#call,rc__); } while (0)
#define NC(call)  do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) \
    { fprintf(stderr,"%s: rank %d: %s: %s\n",Prog,R->rank,#call,ncclGetErrorString(r__)); exit (1); } } while (0)

/* first line start at or after pos */
static int64_t line_cut(int fd, int64_t pos, int64_t fsize)
{ char buf[65536];
  if (pos <= 0) return 0;
  if (pos >= fsize) return fsize;
  if (pread(fd,buf,1,pos-1) == 1 && buf[0] == '\n') return pos;
  while (pos < fsize)
    { ssize_t got = pread(fd,buf,sizeof(buf),pos);
      if (got <= 0) break;
      char *nl = memchr(buf,'\n',(size_t) got);
      if (nl != NULL) return pos + (nl - buf) + 1;
      pos += got;
    }
  return fsize;
}

/* file bytes [from, to) -> device, through two pinned buffers: the pread of a chunk overlaps the copy
   of the chunk before it */
static void load(dx_ctx *ctx, Rank *R, uint8_t *stage[2], uint8_t *d_dst, int64_t from, int64_t to)
{ int k = 0;
  while (from < to)
    { size_t want = (size_t) (to - from) < STAGE ? (size_t) (to - from) : STAGE, have = 0;
      while (have < want)
        { ssize_t got = pread(R->fd_in,stage[k] + have,want - have,from + (int64_t) have);
          if (got <= 0) { fprintf(stderr,"%s: System error, read failed!\n",Prog); exit (2); }
          have += (size_t) got;
        }
      DXC(dx_sync(ctx));                       /* the copy out of the OTHER buffer ran meanwhile */
      DXC(dx_h2d(ctx,d_dst,stage[k],want));
      d_dst += want; from += (int64_t) want; k ^= 1;
    }
  DXC(dx_sync(ctx));
}

/* all-gather / broadcast / all-reduce of a few 64-bit words through small device buffers */
static void gather64(dx_ctx *ctx, Rank *R, int64_t *d_buf, const int64_t *mine, int cnt, int64_t *all)
{ int W = R->o->world;
  DXC(dx_h2d(ctx,d_buf,mine,(size_t) cnt*8));
  NC(ncclAllGather(d_buf,d_buf + cnt,(size_t) cnt,ncclInt64,R->comm,(cudaStream_t) dx_stream(ctx)));
  DXC(dx_d2h(ctx,all,d_buf + cnt,(size_t) cnt*8*W));
  DXC(dx_sync(ctx));
}
static void bcast64(dx_ctx *ctx, Rank *R, int64_t *d_buf, int64_t *val, int cnt, int root)
{ if (R->rank == root) DXC(dx_h2d(ctx,d_buf,val,(size_t) cnt*8));
  NC(ncclBroadcast(d_buf,d_buf,(size_t) cnt,ncclInt64,root,R->comm,(cudaStream_t) dx_stream(ctx)));
  DXC(dx_d2h(ctx,val,d_buf,(size_t) cnt*8));
  DXC(dx_sync(ctx));
}
static void allsum64(dx_ctx *ctx, Rank *R, int64_t *d_buf, uint64_t *val, int cnt)
{ DXC(dx_h2d(ctx,d_buf,val,(size_t) cnt*8));
  NC(ncclAllReduce(d_buf,d_buf,(size_t) cnt,ncclUint64,ncclSum,R->comm,(cudaStream_t) dx_stream(ctx)));
  DXC(dx_d2h(ctx,val,d_buf,(size_t) cnt*8));
  DXC(dx_sync(ctx));
}

static void *rank_main(void *arg)
{ Rank *R = (Rank *) arg;
  const Opts *o = R->o;
  const int W = o->world, r = R->rank;
  dx_ctx *ctx = NULL;
  { int rc = dx_open(r,&ctx); if (rc != DX_OK) fail(NULL,r,"dx_open (no CUDA device? there is no CPU fallback)",rc); }
  uint8_t *stage[2];
  stage[0] = (uint8_t *) dx_pinned_alloc(ctx,STAGE); stage[1] = (uint8_t *) dx_pinned_alloc(ctx,STAGE);
  int64_t *d_x = (int64_t *) dx_device_alloc(ctx,(size_t) (W + 1)*2048*8);
  if (!stage[0] || !stage[1] || !d_x) fail(ctx,r,"allocation",DX_E_NOMEM);
  int64_t all[MAXGPU*4], mine[4];
  double t0 = now();

  /* ---- 1. my byte range, its newline count, the first entry start inside it ------------------- */
  const int64_t lo = line_cut(R->fd_in,R->fsize*r/W,R->fsize), hi = line_cut(R->fd_in,R->fsize*(r+1)/W,R->fsize);
  const size_t nraw = (size_t) (hi - lo);
  uint8_t *d_raw = (uint8_t *) dx_device_alloc(ctx,nraw + 64);
  if (!d_raw) fail(ctx,r,"device allocation",DX_E_NOMEM);
  load(ctx,R,stage,d_raw,lo,hi);
  int64_t nl = 0, off = 0;
  DXC(dx_text_lines_dev(ctx,d_raw,nraw,0,&nl,NULL));
  mine[0] = nl;
  gather64(ctx,R,d_x,mine,1,all);
  int64_t first_line = 0, total_lines = 0;
  for (int q = 0; q < W; q++) { if (q < r) first_line += all[q]; total_lines += all[q]; }
  if (total_lines % 6 != 0)
    { if (r == 0) fprintf(stderr,"%s: Line %lld: incomplete last entry of .quiv file\n",Prog,(long long) total_lines + 1);
      exit (1);
    }
  const int64_t skip = (6 - first_line % 6) % 6;
  DXC(dx_text_lines_dev(ctx,d_raw,nraw,skip,&nl,&off));
  mine[0] = (nraw > 0 && off >= 0 && off < (int64_t) nraw) ? lo + off : -1;
  gather64(ctx,R,d_x,mine,1,all);
  int64_t S[MAXGPU+1];
  S[W] = R->fsize;
  for (int q = W-1; q >= 0; q--) S[q] = (all[q] >= 0) ? all[q] : S[q+1];   /* no entry start: empty shard */
  S[0] = 0;
  const int64_t s0 = S[r], s1 = S[r+1];
  const size_t n = (size_t) (s1 - s0);
  if (n > nraw + SLACK) fail(ctx,r,"an entry longer than 96 MB",DX_E_TOOLONG);

  /* ---- 2. the shard, 16-byte aligned on the device: what I hold, plus my last entry's end ------ */
  uint8_t *d_text = (uint8_t *) dx_device_alloc(ctx,n + 64);
  if (!d_text) fail(ctx,r,"device allocation",DX_E_NOMEM);
  { const int64_t have_to = (s1 < hi) ? s1 : hi;
    if (have_to > s0) DXC(dx_d2d(ctx,d_text,d_raw + (s0 - lo),(size_t) (have_to - s0)));
    DXC(dx_sync(ctx));
    if (s1 > hi) load(ctx,R,stage,d_text + (hi > s0 ? hi - s0 : 0),(hi > s0 ? hi : s0),s1);
  }
  dx_device_free(ctx,d_raw);
  R->t_read = now() - t0; t0 = now();

  /* ---- 3. statistics: in file order while the run characters are open, all at once afterwards ---- */
  dx_qv_stats *st = (dx_qv_stats *) calloc(1,sizeof(dx_qv_stats));
  dx_qv_carry *cy = (dx_qv_carry *) calloc(1,sizeof(dx_qv_carry));
  int64_t *cw = (int64_t *) malloc((3 + 256)*8);            /* delchar, subchar, totchar, sub[256] */
  int scanned = 0;
  cy->delchar = cy->subchar = -1;
  for (int root = 0; root < W; root++)
    { if (r == root && !scanned)
        { DXC(dx_qv_scan_dev(ctx,d_text,n,root == 0 ? NULL : cy,st));
          scanned = 1;
          cw[0] = st->delchar; cw[1] = st->subchar; cw[2] = (int64_t) (cy->totchar + st->totchar);
          for (int k = 0; k < 256; k++) cw[3+k] = (int64_t) st->sub_prefix[k];
        }
      bcast64(ctx,R,d_x,cw,3 + 256,root);
      cy->delchar = (int32_t) cw[0]; cy->subchar = (int32_t) cw[1]; cy->totchar = (uint64_t) cw[2];
      for (int k = 0; k < 256; k++) cy->sub[k] = (uint64_t) cw[3+k];
      if (cy->delchar >= 0 && cy->subchar >= 0)               /* fixed: nobody has to wait any longer */
        { if (!scanned) { DXC(dx_qv_scan_dev(ctx,d_text,n,cy,st)); scanned = 1; }
          break;
        }
    }
  int32_t lastw = 0;
  DXC(dx_qv_last_well(ctx,&lastw));
  uint64_t *sum = (uint64_t *) malloc((6*256 + 2)*8);
  memcpy(sum,st->hist,6*256*8);
  sum[6*256] = st->totchar; sum[6*256+1] = (uint64_t) st->nentries;
  allsum64(ctx,R,d_x,sum,6*256 + 2);
  dx_qv_stats *tot = (dx_qv_stats *) calloc(1,sizeof(dx_qv_stats));
  memcpy(tot->hist,sum,6*256*8);
  tot->totchar = sum[6*256]; tot->nentries = (int64_t) sum[6*256+1];
  tot->delchar = cy->delchar; tot->subchar = cy->subchar;     /* the characters in force at the file's end */
  dx_qv_coding *cd = (dx_qv_coding *) calloc(1,sizeof(dx_qv_coding));
  DXC(dx_qv_make_coding(tot,o->lossy,cd));
  R->t_scan = now() - t0; t0 = now();

  /* ---- 4. header (the same bytes on every rank), hand-off wells, encode, sizes, write ------------- */
  uint8_t *hdr = (uint8_t *) malloc(200000);
  size_t hlen = 0;
  { char prefix[4096]; uint8_t head[4096];
    size_t got = (size_t) pread(R->fd_in,head,sizeof(head)-1,0), pl = 0;
    while (pl < got && head[pl] != '/' && head[pl] != '\n') pl++;
    if (pl >= sizeof(prefix)) pl = sizeof(prefix)-1;
    memcpy(prefix,head,pl); prefix[pl] = 0;
    hdr[0] = 0xaa; hdr[1] = 0x55;                            /* the 0x55aa key, native byte order */
    DXC(dx_qv_write_coding(cd,prefix,(int) pl,hdr + 2,200000 - 2,&hlen));
    hlen += 2;
  }
  mine[0] = st->nentries; mine[1] = lastw;
  gather64(ctx,R,d_x,mine,2,all);
  int32_t lwell_in = 0;
  for (int q = r-1; q >= 0; q--) if (all[2*q] > 0) { lwell_in = (int32_t) all[2*q+1]; break; }
  size_t cap = n/2 + ((size_t) 1 << 20), m = 0;
  uint8_t *d_out = (uint8_t *) dx_device_alloc(ctx,cap);
  if (!d_out) fail(ctx,r,"device allocation",DX_E_NOMEM);
  int rc = dx_qv_encode_dev(ctx,d_text,n,cd,o->lossy,lwell_in,d_out,cap,&m,NULL,NULL,0);
  if (rc == DX_E_CAP)                                         /* worst case: every symbol escaped (24 bits) */
    { dx_device_free(ctx,d_out);
      cap = 3*n + 200000;
      if ((d_out = (uint8_t *) dx_device_alloc(ctx,cap)) == NULL) fail(ctx,r,"device allocation",DX_E_NOMEM);
      rc = dx_qv_encode_dev(ctx,d_text,n,cd,o->lossy,lwell_in,d_out,cap,&m,NULL,NULL,0);
    }
  if (rc != DX_OK) fail(ctx,r,"dx_qv_encode_dev",rc);
  R->t_code = now() - t0; t0 = now();
  mine[0] = (int64_t) m;
  gather64(ctx,R,d_x,mine,1,all);
  int64_t base = (int64_t) hlen;
  for (int q = 0; q < r; q++) base += all[q];
  if (r == 0 && pwrite(R->fd_out,hdr,hlen,0) != (ssize_t) hlen)
    { fprintf(stderr,"%s: System error, write failed!\n",Prog); exit (2); }
  for (size_t at = 0; at < m; at += STAGE)
    { const size_t len = (m - at < STAGE) ? m - at : STAGE;
      DXC(dx_d2h(ctx,stage[0],d_out + at,len));
      DXC(dx_sync(ctx));
      if (pwrite(R->fd_out,stage[0],len,base + (int64_t) at) != (ssize_t) len)
        { fprintf(stderr,"%s: System error, write failed!\n",Prog); exit (2); }
    }
  R->t_write = now() - t0;
  R->nbytes = (int64_t) n; R->nent = st->nentries; R->obytes = (int64_t) m;

  /* ---- 5. -c: decode my shard again ([header][my entries], first well against lwell_in) ---------- */
  if (o->check && n > 0)
    { uint8_t *d_img = (uint8_t *) dx_device_alloc(ctx,hlen + m + 64);
      uint8_t *d_back = (uint8_t *) dx_device_alloc(ctx,n + 4096);
      size_t k = 0;
      if (!d_img || !d_back) fail(ctx,r,"device allocation (-c)",DX_E_NOMEM);
      DXC(dx_h2d(ctx,d_img,hdr,hlen));
      DXC(dx_d2d(ctx,d_img + hlen,d_out,m));
      DXC(dx_undexqv_dev(ctx,d_img,hlen + m,0,d_back,n + 4096,&k,NULL,0,lwell_in));
      if (k != n) { fprintf(stderr,"%s: rank %d: -c: decoded %zu bytes, shard has %zu\n",Prog,r,k,n); exit (1); }
      for (size_t at = 0; at < n; at += STAGE)
        { const size_t len = (n - at < STAGE) ? n - at : STAGE;
          size_t have = 0;
          DXC(dx_d2h(ctx,stage[0],d_back + at,len));
          while (have < len)
            { ssize_t got = pread(R->fd_in,stage[1] + have,len - have,s0 + (int64_t) (at + have));
              if (got <= 0) { fprintf(stderr,"%s: System error, read failed!\n",Prog); exit (2); }
              have += (size_t) got;
            }
          DXC(dx_sync(ctx));
          if (memcmp(stage[0],stage[1],len) != 0)
            { fprintf(stderr,"%s: rank %d: -c: decoded shard differs from the file near byte %lld\n",
                      Prog,r,(long long) (s0 + (int64_t) at));
              exit (1);
            }
        }
      dx_device_free(ctx,d_img); dx_device_free(ctx,d_back);
    }
  dx_device_free(ctx,d_out); dx_device_free(ctx,d_text); dx_device_free(ctx,d_x);
  dx_pinned_free(ctx,stage[0]); dx_pinned_free(ctx,stage[1]);
  free(st); free(cy); free(cw); free(sum); free(tot); free(cd); free(hdr);
  dx_close(ctx);
  return NULL;
}

static char *file_name(const char *arg, const char *strip, const char *ext)
{ size_t la = strlen(arg), ls = strlen(strip);
  char *out = (char *) malloc(la + strlen(ext) + 16);
  strcpy(out,arg);
  if (la > ls && strcasecmp(arg+la-ls,strip) == 0) out[la-ls] = '\0';
  strcat(out,ext);
  return out;
}

int main(int argc, char *argv[])
{ Opts o = { 0, 0, 0, 0, 0 };
  int i, j = 1, k;
  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-')
      { if (argv[i][1] == 'g') { o.world = atoi(argv[i]+2); continue; }
        for (k = 1; argv[i][k] != '\0'; k++)
          switch (argv[i][k])
            { case 'v': o.verbose = 1; break;
              case 'k': o.keep = 1; break;
              case 'l': o.lossy = 1; break;
              case 'c': o.check = 1; break;
              default:
                fprintf(stderr,"%s: -%c is an illegal option\n",Prog,argv[i][k]); exit (1);
            }
      }
    else
      argv[j++] = argv[i];
  argc = j;
  if (argc <= 1 || o.world < 1 || o.world > MAXGPU)
    { fprintf(stderr,"Usage: %s [-vklc] -g<gpus(1..%d)> <path:quiva> ...\n\n",Prog,MAXGPU);
      fprintf(stderr,"      -k: do *not* remove the .quiva file on completion.\n");
      fprintf(stderr,"      -l: use lossy compression (not recommended).\n");
      fprintf(stderr,"      -c: decode every shard again and compare it with the file.\n");
      exit (1);
    }
  ncclComm_t comms[MAXGPU];
  int devs[MAXGPU];
  for (k = 0; k < o.world; k++) devs[k] = k;
  { ncclResult_t r = ncclCommInitAll(comms,o.world,devs);
    if (r != ncclSuccess) { fprintf(stderr,"%s: ncclCommInitAll: %s\n",Prog,ncclGetErrorString(r)); exit (1); }
  }
  for (i = 1; i < argc; i++)
    { char *src = file_name(argv[i],".quiva",".quiva"), *dst = file_name(argv[i],".quiva",".dexqv");
      struct stat sb;
      Rank R[MAXGPU];
      pthread_t th[MAXGPU];
      int fd_in = open(src,O_RDONLY), fd_out;
      double t0 = now();
      if (fd_in < 0 || fstat(fd_in,&sb) != 0)
        { fprintf(stderr,"%s: Cannot open %s for 'r'\n",Prog,src); exit (1); }
      if ((fd_out = open(dst,O_WRONLY|O_CREAT|O_TRUNC,0644)) < 0)
        { fprintf(stderr,"%s: Cannot open %s for 'w'\n",Prog,dst); exit (1); }
      if (o.verbose) { fprintf(stderr,"Processing '%s' on %d GPUs ...\n",src,o.world); fflush(stderr); }
      for (k = 0; k < o.world; k++)
        { memset(&R[k],0,sizeof(Rank));
          R[k].rank = k; R[k].o = &o; R[k].comm = comms[k]; R[k].fd_in = fd_in; R[k].fd_out = fd_out;
          R[k].fsize = (int64_t) sb.st_size;
          if (pthread_create(&th[k],NULL,rank_main,&R[k]) != 0) { fprintf(stderr,"%s: pthread_create failed\n",Prog); exit (1); }
        }
      for (k = 0; k < o.world; k++) pthread_join(th[k],NULL);
      close(fd_in);
      if (close(fd_out) != 0) { fprintf(stderr,"%s: System error, write failed!\n",Prog); exit (2); }
      if (o.verbose)
        { double tr = 0, ts = 0, tc = 0, tw = 0; int64_t ne = 0, ob = 0;
          for (k = 0; k < o.world; k++)
            { if (R[k].t_read > tr)  tr = R[k].t_read;
              if (R[k].t_scan > ts)  ts = R[k].t_scan;
              if (R[k].t_code > tc)  tc = R[k].t_code;
              if (R[k].t_write > tw) tw = R[k].t_write;
              ne += R[k].nent; ob += R[k].obytes;
            }
          fprintf(stderr,"  %lld bytes, %lld entries -> %lld bytes; max over ranks: read+H2D %.3f s, "
                         "scan+exchange %.3f s, encode %.3f s, D2H+write %.3f s; wall %.3f s\n",
                  (long long) sb.st_size,(long long) ne,(long long) ob,tr,ts,tc,tw,now() - t0);
          for (k = 0; k < o.world; k++)
            fprintf(stderr,"  rank %d: %lld bytes, %lld entries -> %lld bytes\n",k,(long long) R[k].nbytes,
                    (long long) R[k].nent,(long long) R[k].obytes);
        }
      if (!o.keep) unlink(src);
      free(src); free(dst);
      if (o.verbose) { fprintf(stderr,"Done\n"); fflush(stderr); }
    }
  for (k = 0; k < o.world; k++) ncclCommDestroy(comms[k]);
  return 0;
}
