// dx_compat.cu -- libdexcompat.so: the reference's own QV.h / DB.h entry points over libdexb200.so.
//
// the reference tools call the codec one read at a time through FILE* streams
// (QVcoding_Scan, Create_QVcoding, Write_QVcoding, Compress_Next_QVentry, Read_QVcoding,
// Uncompress_Next_QVentry of QV.h:48-97; Compress_Read, Uncompress_Read, Number_Read, Lower_Read,
// Upper_Read, Number_Arrow, Letter_Arrow, Change_Read of DB.h:255-267; the Malloc/Fopen/PathTo/Root/
// Catenate helpers of DB.h:235-247).  This file exports exactly those C symbols, so that the
// reference's mains (dexqv.c, undexqv.c, dexta.c, ...) compile against their own headers and link
// here instead of against DB.c + QV.c.  All codec work runs on the GPU through the batch ABI of
// include/dexb200.h: the per-entry calls are served by "capture at scan, replay at compress" --
// the reference's call pattern (scan all entries, create the coding, compress the same entries in the
// same order; read the coding, then decode entries) is what makes that exact.  Nothing here
// encodes or decodes on the host; line reading and path handling are host I/O, as in the reference.
//
// One process-wide context on device 0, not thread safe -- like the statics of QV.c:35-38, 733-736,
// 860-862 (SURVEY 8b "Threading").

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <stdint.h>
#include <limits.h>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "dexb200.h"

typedef long long int64;

// QV.h:31-42 (layout is part of the interface: callers read delChar/subChar/flip and own prefix)
typedef struct
  { void *delScheme, *insScheme, *mrgScheme, *subScheme, *dRunScheme, *sRunScheme;
    int   delChar, subChar, flip;
    char *prefix;
  } QVcoding;

extern "C" { extern char *Prog_Name; }

// Error convention (DB.h:28-47, QV.h:20-27; SURVEY 8b "Errors").  Batch build (libdexcompat.so):
// message to stderr, exit(1) (exit(2) for a failed read).  -DINTERACTIVE build (libdexcompat_i.so):
// the message goes to the exported Ebuffer[1000] (DB.c:42) and the routine returns its documented
// error value -- NULL for pointers, -1 / -2 / 1 for the integer routines.  Helpers below the entry
// points report through fatal(); in the interactive build that unwinds to the entry point by a C++
// exception, which never crosses the C boundary.
#ifdef INTERACTIVE
extern "C" { char Ebuffer[1000]; }
#define DXC_MSG(...)        snprintf(Ebuffer,sizeof(Ebuffer),__VA_ARGS__)
#define DXC_EXIT(code,ret)  return ret
#define DXC_TRY             try
#define DXC_CATCH(ret)      catch (const ShimError &) { return ret; }
#define DXC_CATCH_VOID      catch (const ShimError &) { return; }
#else
#define DXC_MSG(...)        fprintf(stderr,__VA_ARGS__)
#define DXC_EXIT(code,ret)  exit (code)
#define DXC_TRY
#define DXC_CATCH(ret)
#define DXC_CATCH_VOID
#endif

namespace {

struct ShimError {};

// ---- tiny kernels for the per-read DB.h calls ---------------------------------------------------------
__global__ void k_map(uint8_t *s, int n, const uint8_t *table)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) s[i] = table[s[i]];
}

__global__ void k_pack_numeric(const uint8_t *s, int len, uint8_t *out)          // DB.c:319-338
{ const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (4*j >= len) return;
  uint32_t v = 0;
  for (int k = 0; k < 4; k++)
    v = (v << 2) | ((4*j + k < len) ? (s[4*j + k] & 3u) : 0u);
  out[j] = (uint8_t) v;
}

__global__ void k_unpack_numeric(const uint8_t *in, int clen, uint8_t *out)        // DB.c:342-363
{ const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= clen) return;
  const uint32_t b = in[j];
  out[4*j] = (uint8_t) (b >> 6); out[4*j+1] = (uint8_t) ((b >> 4) & 3u);
  out[4*j+2] = (uint8_t) ((b >> 2) & 3u); out[4*j+3] = (uint8_t) (b & 3u);
}

struct Shim
{ dx_ctx *ctx = NULL;
  uint8_t *d_a = NULL, *d_b = NULL, *d_tab = NULL; size_t cap_a = 0, cap_b = 0;
  // encoder side
  std::vector<uint8_t> text;            // captured .quiva text (QVcoding_Scan / QVcoding_Scan1)
  uint8_t *d_text = NULL; size_t d_text_cap = 0;
  bool scanned = false, encoded = false;
  int64_t scan1_entries = 0;
  dx_qv_stats stats;
  std::vector<uint8_t> image;           // all entries, encoded
  std::vector<int64_t> eoff;
  int64_t next = 0;
  // decoder side
  std::vector<uint8_t> dimg;            // key + coding header + entries
  int64_t dpos0 = 0;                    // file offset of dimg[dkey]
  int     dkey = 2;                     // key bytes put in front of the captured image (0: old layout)
  bool decoded = false, undecodable = false;
  std::vector<uint8_t> dtext;
  std::vector<dx_index_row> index;
  // line reader (QV.c:733-798)
  char *line = NULL; int rmax = 0; int nline = 0;
};
Shim S;

void fatal(const char *what)
{ DXC_MSG("%s: %s%s%s\n",Prog_Name ? Prog_Name : "dexcompat",what,S.ctx ? ": " : "",
          S.ctx ? dx_strerror(S.ctx) : "");
#ifdef INTERACTIVE
  throw ShimError();
#else
  exit (1);
#endif
}

dx_ctx *gpu()
{ if (S.ctx == NULL && dx_open(0,&S.ctx) != DX_OK)
    { S.ctx = NULL; fatal("no usable CUDA device (there is no CPU fallback)"); }
  return S.ctx;
}

void need(uint8_t **p, size_t *cap, size_t n)
{ if (*cap >= n + 64) return;
  if (*p) dx_device_free(gpu(),*p);
  *cap = n + n/2 + 4096;
  *p = (uint8_t *) dx_device_alloc(gpu(),*cap);
  if (*p == NULL) fatal("device allocation failed");
}

// in-place byte map of s[0..n) on the GPU
void map_bytes(char *s, int n, const uint8_t table[256])
{ if (n <= 0) return;
  dx_ctx *c = gpu();
  need(&S.d_a,&S.cap_a,(size_t) n);
  if (S.d_tab == NULL) S.d_tab = (uint8_t *) dx_device_alloc(c,256);
  cudaStream_t st = (cudaStream_t) dx_stream(c);
  dx_h2d(c,S.d_tab,table,256);
  dx_h2d(c,S.d_a,s,(size_t) n);
  k_map<<<(n + 255)/256,256,0,st>>>(S.d_a,n,S.d_tab);
  dx_d2h(c,s,S.d_a,(size_t) n);
  if (dx_sync(c) != DX_OK) fatal("GPU error");
}

struct Tables
{ uint8_t number[256], lower[256], upper[256], narrow[256], larrow[256], change[256];
  Tables()
  { for (int i = 0; i < 256; i++)
      { number[i] = 0; lower[i] = upper[i] = larrow[i] = (uint8_t) i; narrow[i] = 3; change[i] = (uint8_t) i; }
    number['c'] = number['C'] = 1; number['g'] = number['G'] = 2; number['t'] = number['T'] = 3;   // DB.c:394-411
    narrow['1'] = 0; narrow['2'] = 1; narrow['3'] = 2; narrow['G'] = 2; narrow['4'] = 3;           // DB.c:419-436
    for (int i = 0; i < 4; i++)
      { lower[i] = (uint8_t) "acgt"[i]; upper[i] = (uint8_t) "ACGT"[i]; larrow[i] = (uint8_t) "1234"[i]; }
    for (int i = 'a'; i <= 'z'; i++) { change[i] = (uint8_t) (i - 32); change[i - 32] = (uint8_t) i; }
  }
} T;

dx_qv_coding *dxc(QVcoding *c) { return (dx_qv_coding *) c->delScheme; }

void publish(QVcoding *out, const dx_qv_coding *cd)
{ dx_qv_coding *copy = (dx_qv_coding *) malloc(sizeof(dx_qv_coding));
  if (copy == NULL) fatal("Out of memory (coding)");
  *copy = *cd;
  out->delScheme = copy;                                  // the other five only signal "present"
  out->insScheme = out->mrgScheme = out->subScheme = copy;
  out->dRunScheme = (cd->delchar >= 0) ? copy : NULL;
  out->sRunScheme = (cd->subchar >= 0) ? copy : NULL;
  out->delChar = cd->delchar; out->subChar = cd->subchar; out->flip = cd->flip;
  out->prefix = NULL;
}

void upload_text()
{ dx_ctx *c = gpu();
  need(&S.d_text,&S.d_text_cap,S.text.size());
  if (!S.text.empty() && dx_h2d(c,S.d_text,S.text.data(),S.text.size()) != DX_OK) fatal("copy to the device failed");
  dx_sync(c);
}

void run_scan()
{ upload_text();
  if (dx_qv_scan_dev(gpu(),S.d_text,S.text.size(),NULL,&S.stats) != DX_OK) fatal("QVcoding_Scan");
  S.scanned = true; S.encoded = false; S.next = 0;
}

void run_encode(QVcoding *coding, int lossy)
{ dx_ctx *c = gpu();
  const size_t cap = 3*S.text.size() + 200000;
  need(&S.d_b,&S.cap_b,cap);
  size_t m = 0; int32_t lastw = 0;
  S.eoff.assign((size_t) S.stats.nentries + 1,0);
  if (dx_qv_encode_dev(c,S.d_text,S.text.size(),dxc(coding),lossy,0,S.d_b,cap,&m,&lastw,S.eoff.data(),
                       S.stats.nentries) != DX_OK) fatal("Compress_Next_QVentry");
  S.image.resize(m);
  if (m && dx_d2h(c,S.image.data(),S.d_b,m) != DX_OK) fatal("copy from the device failed");
  dx_sync(c);
  S.encoded = true; S.next = 0;
}

// write the streams of captured entry k (its well-delta bytes and beg/end/qv are the caller's job)
void emit_entry(FILE *output)
{ if (S.next >= (int64_t) S.eoff.size() - 1) fatal("Compress_Next_QVentry called more often than entries were scanned");
  const uint8_t *p = S.image.data() + S.eoff[(size_t) S.next], *e = S.image.data() + S.eoff[(size_t) S.next + 1];
  while (p < e && *p == 0xff) p++;
  p += 13;
  if (p > e) fatal("internal error: short entry");
  if (e > p) fwrite(p,1,(size_t) (e - p),output);
  S.next++;
}

}  // namespace

extern "C" {

char *Prog_Name = NULL;

// ---- DB.h:235-247 utilities (host, not on the hot path) -------------------------------------------------
void *Malloc(int64 size, char *mesg)
{ void *p = malloc((size_t) size);
  if (p == NULL) DXC_MSG(mesg ? "%s: Out of memory (%s)\n" : "%s: Out of memory\n",Prog_Name,mesg);
  return p;
}

void *Realloc(void *p, int64 size, char *mesg)
{ p = realloc(p,(size_t) (size > 0 ? size : 1));
  if (p == NULL) DXC_MSG(mesg ? "%s: Out of memory (%s)\n" : "%s: Out of memory\n",Prog_Name,mesg);
  return p;
}

char *Strdup(char *string, char *mesg)
{ if (string == NULL) return NULL;
  char *s = strdup(string);
  if (s == NULL) DXC_MSG(mesg ? "%s: Out of memory (%s)\n" : "%s: Out of memory\n",Prog_Name,mesg);
  return s;
}

FILE *Fopen(char *path, char *mode)
{ if (path == NULL || mode == NULL) return NULL;
  FILE *f = fopen(path,mode);
  if (f == NULL) DXC_MSG("%s: Cannot open %s for '%s'\n",Prog_Name,path,mode);
  return f;
}

char *PathTo(char *path)
{ if (path == NULL) return NULL;
  const char *sl = strrchr(path,'/');
  if (sl == NULL) return Strdup((char *) ".",(char *) "Allocating default path");
  char *out = (char *) Malloc((int64) (sl - path) + 1,(char *) "Extracting path from");
  if (out != NULL) { memcpy(out,path,(size_t) (sl - path)); out[sl - path] = '\0'; }
  return out;
}

char *Root(char *path, char *suffix)
{ if (path == NULL) return NULL;
  const char *base = strrchr(path,'/');
  base = base ? base + 1 : path;
  size_t keep = strlen(base);
  if (suffix == NULL)
    { const char *dot = strchr(base,'.');
      if (dot) keep = (size_t) (dot - base);
    }
  else
    { const size_t ls = strlen(suffix);
      if (keep > ls && strcasecmp(base + keep - ls,suffix) == 0) keep -= ls;
    }
  char *out = (char *) Malloc((int64) keep + 1,(char *) "Extracting root from");
  if (out != NULL) { memcpy(out,base,keep); out[keep] = '\0'; }
  return out;
}

char *Catenate(char *path, char *sep, char *root, char *suffix)
{ static std::vector<char> buf;
  if (path == NULL || sep == NULL || root == NULL || suffix == NULL) return NULL;
  buf.resize(strlen(path) + strlen(sep) + strlen(root) + strlen(suffix) + 1);
  sprintf(buf.data(),"%s%s%s%s",path,sep,root,suffix);
  return buf.data();
}

char *Numbered_Suffix(char *left, int num, char *right)
{ static std::vector<char> buf;
  if (left == NULL || right == NULL) return NULL;
  buf.resize(strlen(left) + strlen(right) + 48);
  sprintf(buf.data(),"%s%d%s",left,num,right);
  return buf.data();
}

// ---- DB.h:255-267: the 2-bit codec and alphabet maps, one read at a time, on the GPU ---------------------
void Number_Read(char *s)
{ const int n = (int) strlen(s);
  DXC_TRY { map_bytes(s,n,T.number); } DXC_CATCH_VOID
  s[n] = 4;
}

void Number_Arrow(char *s)
{ const int n = (int) strlen(s);
  DXC_TRY { map_bytes(s,n,T.narrow); } DXC_CATCH_VOID
  s[n] = 4;
}

static void letters(char *s, const uint8_t *table)
{ int n = 0;
  while (s[n] != 4) n++;
  DXC_TRY { map_bytes(s,n,table); } DXC_CATCH_VOID
  s[n] = '\0';
}

void Lower_Read(char *s)   { letters(s,T.lower); }
void Upper_Read(char *s)   { letters(s,T.upper); }
void Letter_Arrow(char *s) { letters(s,T.larrow); }
void Change_Read(char *s)  { DXC_TRY { map_bytes(s,(int) strlen(s),T.change); } DXC_CATCH_VOID }

void Compress_Read(int len, char *s)
{ if (len <= 0) { if (len == 0) s[0] = 0; return; }
  DXC_TRY {
  dx_ctx *c = gpu();
  const int clen = (len + 3) >> 2;
  need(&S.d_a,&S.cap_a,(size_t) len); need(&S.d_b,&S.cap_b,(size_t) clen);
  cudaStream_t st = (cudaStream_t) dx_stream(c);
  dx_h2d(c,S.d_a,s,(size_t) len);
  k_pack_numeric<<<(clen + 255)/256,256,0,st>>>(S.d_a,len,S.d_b);
  dx_d2h(c,s,S.d_b,(size_t) clen);
  if (dx_sync(c) != DX_OK) fatal("GPU error");
  if (clen < len) s[len] = 0;                     // what DB.c:329-337 leaves behind
  } DXC_CATCH_VOID
}

void Uncompress_Read(int len, char *s)
{ const int clen = (len + 3) >> 2;
  if (clen > 0)
    DXC_TRY
    { dx_ctx *c = gpu();
      need(&S.d_a,&S.cap_a,(size_t) clen); need(&S.d_b,&S.cap_b,(size_t) 4*clen);
      cudaStream_t st = (cudaStream_t) dx_stream(c);
      dx_h2d(c,S.d_a,s,(size_t) clen);
      k_unpack_numeric<<<(clen + 255)/256,256,0,st>>>(S.d_a,clen,S.d_b);
      dx_d2h(c,s,S.d_b,(size_t) 4*clen);          // like DB.c:352-362 this may write up to 3 bytes past len
      if (dx_sync(c) != DX_OK) fatal("GPU error");
    }
    DXC_CATCH_VOID
  s[len] = 4;
}

// ---- QV.h: line reader (host I/O, QV.c:733-798) --------------------------------------------------------------
void Set_QV_Line(int line) { S.nline = line; }
int  Get_QV_Line()         { return S.nline; }
char *QVentry()            { return S.line; }

int Read_Lines(FILE *input, int nlines)
{ if (S.line == NULL)
    { S.rmax = 50000;
      S.line = (char *) malloc((size_t) 5*S.rmax);
      if (S.line == NULL) { DXC_MSG("%s: Out of memory (Allocating QV entry read buffer)\n",Prog_Name); DXC_EXIT(1,-2); }
    }
  // Observable behaviour of QV.c:751-798, kept to the letter: the line counter moves BEFORE every read
  // (so it has moved when end of input is met); only the first line may outgrow the buffer, and a
  // first line cut short by end of input is the "no newline" error; any later line whose length
  // (newline included) differs from the first -- also a last line without its newline -- is the
  // "not the same length" error; lines are compared with their newline and the length returned
  // without it.
  S.nline++;
  if (fgets(S.line,S.rmax,input) == NULL) return -1;
  int len = (int) strlen(S.line);
  if (len == 0)                                             // a NUL at the start of a line: the reference reads Read[-1] here
    { DXC_MSG("Line %d: Last line does not end with a newline !\n",S.nline); DXC_EXIT(1,-2); }
  while (S.line[len-1] != '\n')
    { const int nmax = S.rmax + S.rmax/2 + 1000;               // only slot 0 is live here
      char *nl = (char *) malloc((size_t) 5*nmax);
      if (nl == NULL) { DXC_MSG("%s: Out of memory (Reallocating QV entry read buffer)\n",Prog_Name); DXC_EXIT(1,-2); }
      memcpy(nl,S.line,(size_t) len + 1);
      free(S.line); S.line = nl; S.rmax = nmax;
      if (fgets(S.line + len,S.rmax - len,input) == NULL)
        { DXC_MSG("Line %d: Last line does not end with a newline !\n",S.nline);
          DXC_EXIT(1,-2);
        }
      len += (int) strlen(S.line + len);
    }
  for (int i = 1; i < nlines; i++)
    { char *dst = S.line + (size_t) i*S.rmax;
      S.nline++;
      if (fgets(dst,S.rmax,input) == NULL)
        { DXC_MSG("Line %d: incomplete last entry of .quiv file\n",S.nline);
          DXC_EXIT(1,-2);
        }
      if ((int) strlen(dst) != len)
        { DXC_MSG("Line %d: Lines for an entry are not the same length\n",S.nline);
          DXC_EXIT(1,-2);
        }
    }
  return len - 1;
}

// ---- QV.h: statistics, coding (GPU scan; tables on the host exactly like dx_qv_make_coding) ------------
int QVcoding_Scan(FILE *input, int num, FILE *temp)
{ // capture the entries (all that follow, or the next num) and scan them on the GPU
  S.text.clear(); S.scan1_entries = 0;
  const off_t at = ftello(input);
  int64_t got = 0;
  if (num == INT_MAX)
    { char buf[1 << 16]; size_t k;
      while ((k = fread(buf,1,sizeof(buf),input)) > 0) S.text.insert(S.text.end(),buf,buf + k);
    }
  else
    { int c, lines = 0;
      while (lines < 6*(int64_t) num && (c = fgetc(input)) != EOF)
        { S.text.push_back((uint8_t) c);
          if (c == '\n') lines++;
        }
    }
  (void) at;
  if (temp != NULL && !S.text.empty()) fwrite(S.text.data(),1,S.text.size(),temp);
  DXC_TRY { run_scan(); } DXC_CATCH(-1)
  got = S.stats.nentries;
  S.nline += (int) (6*got);
  return (int) got;
}

void QVcoding_Scan1(int rlen, char *del, char *tag, char *ins, char *mrg, char *sub)
{ if (rlen == 0) { S.text.clear(); S.scan1_entries = 0; S.scanned = false; return; }    // QV.c:868-885
  char hdr[64];
  const int hl = sprintf(hdr,"@s/%lld/0_%d RQ=0.0\n",(long long) S.scan1_entries++,rlen);
  S.text.insert(S.text.end(),hdr,hdr + hl);
  const char *l[5] = { del, tag, ins, mrg, sub };
  for (int k = 0; k < 5; k++)
    { S.text.insert(S.text.end(),l[k],l[k] + rlen); S.text.push_back('\n'); }
  S.scanned = false;
}

QVcoding *Create_QVcoding(int lossy)
{ static QVcoding coding;
  DXC_TRY {
  if (!S.scanned) run_scan();
  dx_qv_coding cd;
  if (dx_qv_make_coding(&S.stats,lossy,&cd) != DX_OK)
    { DXC_MSG("%s: a QV stream has fewer than two distinct symbols\n",Prog_Name); DXC_EXIT(1,NULL); }
  publish(&coding,&cd);
  } DXC_CATCH(NULL)
  return &coding;
}

void Write_QVcoding(FILE *output, QVcoding *coding)
{ std::vector<uint8_t> buf(20000 + (coding->prefix ? strlen(coding->prefix) : 0));
  size_t n = 0;
  if (dx_qv_write_coding(dxc(coding),coding->prefix ? coding->prefix : "",
                         coding->prefix ? (int) strlen(coding->prefix) : 0,buf.data(),buf.size(),&n) != DX_OK)
    DXC_TRY { fatal("Write_QVcoding"); } DXC_CATCH_VOID
  fwrite(buf.data(),1,n,output);
}

QVcoding *Read_QVcoding(FILE *input)
{ static QVcoding coding;
  S.dpos0 = (int64_t) ftello(input);
  // undexqv.c:103-110: a caller that found no 0x55aa key rewinds -- position 0 means the OLD layout
  // (the file begins with the coding header, entry fields are uint16); anywhere else the key the
  // caller has already consumed is put back in front of the image
  S.dkey = (S.dpos0 == 0) ? 0 : 2;
  S.dimg.assign((size_t) S.dkey,0);
  if (S.dkey) { S.dimg[0] = 0xaa; S.dimg[1] = 0x55; }
  { char buf[1 << 16]; size_t k;
    while ((k = fread(buf,1,sizeof(buf),input)) > 0) S.dimg.insert(S.dimg.end(),buf,buf + k);
  }
  dx_qv_coding cd;
  std::vector<char> prefix(100001);
  size_t used = 0;
  if (dx_qv_read_coding(S.dimg.data() + S.dkey,S.dimg.size() - (size_t) S.dkey,&cd,prefix.data(),(int) prefix.size(),&used) != DX_OK)
    { DXC_MSG("%s: Could not read the coding scheme (Read_QVcoding)\n",Prog_Name); DXC_EXIT(2,NULL); }
  DXC_TRY { publish(&coding,&cd); } DXC_CATCH(NULL)
  coding.prefix = strdup(prefix.data());
  fseeko(input,(off_t) (S.dpos0 + (int64_t) used),SEEK_SET);
  S.decoded = false; S.undecodable = false;
  return &coding;
}

void Free_QVcoding(QVcoding *coding)
{ if (coding->delScheme) free(coding->delScheme);
  coding->delScheme = coding->insScheme = coding->mrgScheme = coding->subScheme = NULL;
  coding->dRunScheme = coding->sRunScheme = NULL;
  free(coding->prefix);
  coding->prefix = NULL;
}

// ---- QV.h: entries ------------------------------------------------------------------------------------------
int Compress_Next_QVentry(FILE *input, FILE *output, QVcoding *coding, int lossy)
{ const int rlen = Read_Lines(input,5);                      // keep the FILE* where the reference would
  if (rlen < 0) { if (rlen == -1) DXC_MSG("Line %d: incomplete last entry of .quiv file\n",S.nline); DXC_EXIT(1,-1); }
  DXC_TRY
  { if (!S.encoded) run_encode(coding,lossy);
    emit_entry(output);
  }
  DXC_CATCH(-1)
  return rlen;
}

void Compress_Next_QVentry1(int rlen, char *del, char *tag, char *ins, char *mrg, char *sub,
                            FILE *output, QVcoding *coding, int lossy)
{ (void) rlen; (void) del; (void) tag; (void) ins; (void) mrg; (void) sub;
  DXC_TRY
  { if (!S.encoded) run_encode(coding,lossy);
    emit_entry(output);
  }
  DXC_CATCH_VOID
}

// Two ways to serve the call, both leave the FILE* exactly where the reference would (QV.c:1428-1481):
//   * the file Read_QVcoding captured is a .dexqv (undexqv.c's loop): the whole image is decoded once,
//     every call hands out the entry whose streams start at the current position;
//   * anything else -- a Dazzler .qvs (bare streams, several codings in one file, DB.c:2450-2507) read
//     after fseeko(coff) (DB.c:2598-2599), or a position the index does not know: the bytes that follow
//     the position are decoded as ONE entry with the coding the caller passes (dx_qv_load_entries_dev).
static int one_entry(FILE *input, char **entry, QVcoding *coding, int rlen)
{ dx_ctx *c = gpu();
  const off_t pos = ftello(input);
  // worst case: every symbol escaped (16-bit code + 8-bit literal) in four streams, plus the tags
  const size_t want = (size_t) rlen*13 + 256;
  std::vector<uint8_t> buf(want);
  const size_t got = fread(buf.data(),1,want,input);
  need(&S.d_a,&S.cap_a,got + 64);
  const size_t outn = 5*((size_t) rlen + 1);
  need(&S.d_b,&S.cap_b,outn + 64);
  if (got) dx_h2d(c,S.d_a,buf.data(),got);
  int64_t so = 0, eo = 0; int32_t rl = rlen;
  if (dx_qv_load_entries_dev(c,S.d_a,got,dxc(coding),&so,&rl,1,0,S.d_b,S.cap_b,NULL,&eo) != DX_OK)
    { DXC_MSG("%s: Could not read more bits (Decode)\n",Prog_Name ? Prog_Name : "dexcompat"); return 1; }
  std::vector<uint8_t> lines(outn);
  dx_d2h(c,lines.data(),S.d_b,outn);
  if (dx_sync(c) != DX_OK) fatal("GPU error");
  for (int e = 0; e < 5; e++)
    memcpy(entry[e],lines.data() + (size_t) e*((size_t) rlen + 1),(size_t) rlen);
  fseeko(input,pos + (off_t) eo,SEEK_SET);
  return 0;
}

int Uncompress_Next_QVentry(FILE *input, char **entry, QVcoding *coding, int rlen)
{ if (!S.decoded && !S.undecodable && S.dimg.size() > 2)
    DXC_TRY
    { dx_ctx *c = gpu();
      need(&S.d_a,&S.cap_a,S.dimg.size());
      dx_h2d(c,S.d_a,S.dimg.data(),S.dimg.size());
      size_t want = 0, m = 0;
      S.undecodable = true;                                  // until the whole image proves to be a .dexqv
      if (dx_undexqv_size_dev(c,S.d_a,S.dimg.size(),&want) == DX_OK)
        { need(&S.d_b,&S.cap_b,want);
          dx_keep_index(c,1);
          const int rc = dx_undexqv_dev(c,S.d_a,S.dimg.size(),0,S.d_b,S.cap_b,&m,NULL,0,0);
          int64_t cnt = 0;
          if (rc == DX_OK) dx_last_index(c,NULL,0,&cnt);
          if (rc == DX_OK && (cnt > 0 || m == 0))
            { S.index.resize((size_t) cnt);
              dx_last_index(c,S.index.data(),cnt,&cnt);
              S.dtext.resize(m);
              if (m) dx_d2h(c,S.dtext.data(),S.d_b,m);
              dx_sync(c);
              S.decoded = true; S.undecodable = false;
            }
          dx_keep_index(c,0);
        }
    }
    DXC_CATCH(1)
  if (S.decoded)
    { // which entry starts at the current file position
      const int64_t img = (int64_t) ftello(input) - S.dpos0 + S.dkey;
      size_t lo = 0, hi = S.index.size();
      while (lo < hi) { const size_t mid = (lo + hi)/2; if (S.index[mid].stream_off < img) lo = mid + 1; else hi = mid; }
      if (lo < S.index.size() && S.index[lo].stream_off == img && S.index[lo].rlen == rlen)
        { const dx_index_row &r = S.index[lo];
          for (int e = 0; e < 5; e++)
            memcpy(entry[e],S.dtext.data() + r.text_off + (int64_t) e*(rlen + 1),(size_t) rlen);
          fseeko(input,(off_t) (S.dpos0 - S.dkey + r.end_off),SEEK_SET);
          return 0;
        }
    }
  int rc = 1;
  DXC_TRY { rc = one_entry(input,entry,coding,rlen); } DXC_CATCH(1)
  return rc;
}

}  // extern "C"
