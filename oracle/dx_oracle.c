/* TEST INFRASTRUCTURE ONLY -- see dx_oracle.h.  Parity status: PINNED against oracle/_ref
 * (the reference tools compiled from the mounted sources) and tests/golden/.
 *
 * Sequential, buffer-to-buffer restatement of the reference's compression hot path.
 * Citations are file:line into /root/reference.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dx_oracle.h"

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t  i64;

#define LINE_LIMIT 100000       /* dexta.c:21 / dexar.c MAX_BUFFER */
#define CUTOFF     16           /* QV.c:26 HUFF_CUTOFF */

/* ------------------------------------------------------------------------------------------
 *  2-bit codec primitives
 * ---------------------------------------------------------------------------------------- */

static int base_code(int c)     /* DB.c:394-411: a/A 0, c/C 1, g/G 2, t/T 3, anything else 0 */
{ switch (c)
    { case 'c': case 'C': return 1;
      case 'g': case 'G': return 2;
      case 't': case 'T': return 3;
      default:            return 0;
    }
}

static int arrow_code(int c)    /* DB.c:419-436: '1'..'4' -> 0..3, 'G' -> 2, anything else 3 */
{ switch (c)
    { case '1': return 0;
      case '2': return 1;
      case '3': return 2;
      case 'G': return 2;
      default:  return 3;
    }
}

void orc_number_read(char *s)
{ for ( ; *s != '\0'; s++)
    *s = (char) base_code(*s);
  *s = 4;
}

void orc_number_arrow(char *s)
{ for ( ; *s != '\0'; s++)
    *s = (char) arrow_code(*s);
  *s = 4;
}

/* DB.c:319-338.  Four numeric symbols per byte, first symbol in the top two bits, the tail of
 * the last byte zero filled.  In place; s[len] ends up 0, s[len+1], s[len+2] are preserved. */
void orc_compress_read(int len, char *s)
{ char keep1 = s[len+1], keep2 = s[len+2];
  int  i, j;

  s[len] = s[len+1] = s[len+2] = 0;
  for (i = 0, j = 0; i < len; i += 4, j++)
    s[j] = (char) ((s[i] << 6) | (s[i+1] << 4) | (s[i+2] << 2) | s[i+3]);
  s[len+1] = keep1;
  s[len+2] = keep2;
}

/* DB.c:342-363.  Inverse, expanding back to front in place, then the terminator 4 at s[len].
 * Like the reference it expands byte 0 even when len == 0. */
void orc_uncompress_read(int len, char *s)
{ int last = (len-1)/4;
  int j;

  for (j = last; j >= 0; j--)
    { int b = (u8) s[j];
      s[4*j+3] = (char) (b & 3);
      s[4*j+2] = (char) ((b >> 2) & 3);
      s[4*j+1] = (char) ((b >> 4) & 3);
      s[4*j]   = (char) ((b >> 6) & 3);
    }
  s[len] = 4;
}

static void letters(char *s, const char *alpha)     /* DB.c:367-389 */
{ for ( ; *s != 4; s++)
    *s = alpha[(int) *s];
  *s = '\0';
}

void orc_lower_read(char *s)   { letters(s,"acgt"); }
void orc_upper_read(char *s)   { letters(s,"ACGT"); }
void orc_letter_arrow(char *s) { letters(s,"1234"); }

/* ------------------------------------------------------------------------------------------
 *  Output buffer helper
 * ---------------------------------------------------------------------------------------- */

typedef struct { u8 *p; i64 n, cap; int over; } obuf;

static void put(obuf *o, const void *src, i64 k)
{ if (o->n + k > o->cap)
    { o->over = 1; return; }
  memcpy(o->p + o->n, src, (size_t) k);
  o->n += k;
}

static void put_u8 (obuf *o, u8 v)   { put(o,&v,1); }
static void put_u16(obuf *o, u16 v)  { put(o,&v,2); }
static void put_i32(obuf *o, int32_t v) { put(o,&v,4); }

static void put_well(obuf *o, int well, int *lwell)      /* dexta.c:187-194, dexqv.c:128-135 */
{ while (well - *lwell >= 255)
    { put_u8(o,0xff);
      *lwell += 255;
    }
  put_u8(o,(u8) (well - *lwell));
  *lwell = well;
}

/* ------------------------------------------------------------------------------------------
 *  dexta / dexar  (dexta.c:100-205, dexar.c:100-211)
 * ---------------------------------------------------------------------------------------- */

/* fgets(buf,LINE_LIMIT,f) stand-in: line starts at *pos; returns 0 at end of input, 1 for a
 * complete line (length incl. '\n' in *len), -1 if it is unterminated or too long. */
static int take_line(const u8 *t, i64 n, i64 *pos, i64 *len)
{ i64 p = *pos, e;

  if (p >= n) return 0;
  for (e = p; e < n && t[e] != '\n'; e++)
    ;
  if (e >= n || e - p + 1 > LINE_LIMIT-1)
    return -1;
  *len = e - p + 1;
  *pos = e + 1;
  return 1;
}

i64 orc_dexta(const u8 *text, i64 n, int arrow, u8 *out, i64 cap)
{ obuf  o = { out, 0, cap, 0 };
  i64   pos = 0, hs, hl, ls, ll;
  int   r, lwell = 0, have;
  char *hdr, *seq;
  i64   smax = 1<<16;

  hs = 0;
  r  = take_line(text,n,&pos,&hl);
  if (r <= 0) return (r == 0 ? ORC_E_FORMAT : ORC_E_TOOLONG);
  if (text[0] != '>') return ORC_E_FORMAT;
  { const u8 *slash = memchr(text,'/',(size_t) hl);
    if (slash == NULL) return ORC_E_FORMAT;
    put_u16(&o,0x55aa);
    put_i32(&o,(int32_t) (slash-text));
    put(&o,text,slash-text);
  }

  seq  = malloc((size_t) smax+8);
  hdr  = malloc(LINE_LIMIT+8);
  have = 1;
  while (have)
    { int   well, beg, end, qv = 0, x;
      float snr[4];
      u16   cnr[4];
      char *slash;
      i64   rlen = 0;

      memcpy(hdr,text+hs,(size_t) hl);               /* header text incl. '\n' */
      hdr[hl] = '\0';
      slash = strchr(hdr+1,'/');
      if (slash == NULL) { o.n = ORC_E_FORMAT; break; }
      if (arrow)
        { x = sscanf(slash+1,"%d/%d_%d SN=%f,%f,%f,%f\n",&well,&beg,&end,snr,snr+1,snr+2,snr+3);
          if (x != 7) { o.n = ORC_E_FORMAT; break; }
          for (x = 0; x < 4; x++)                    /* dexar.c:159-163 */
            if (snr[x] > 99.99)
              cnr[x] = 9999;
            else
              cnr[x] = (u16) (u32) (snr[x]*100.);
        }
      else
        { x = sscanf(slash+1,"%d/%d_%d RQ=0.%d\n",&well,&beg,&end,&qv);
          if (x < 3) { o.n = ORC_E_FORMAT; break; }
          if (x == 3) qv = 0;
        }

      have = 0;                                      /* gather sequence lines: dexta.c:161-183 */
      while (1)
        { ls = pos;
          r  = take_line(text,n,&pos,&ll);
          if (r < 0) { o.n = ORC_E_TOOLONG; goto done; }
          if (r == 0) break;
          if (text[ls] == '>')
            { hs = ls; hl = ll; have = 1; break; }
          if (rlen + ll > smax)
            { smax = 2*(rlen+ll);
              seq  = realloc(seq,(size_t) smax+8);
            }
          memcpy(seq+rlen,text+ls,(size_t) (ll-1));
          rlen += ll-1;
        }
      seq[rlen] = '\0';

      put_well(&o,well,&lwell);
      put_i32(&o,beg);
      put_i32(&o,end);
      if (arrow)
        put(&o,cnr,8);
      else
        put_i32(&o,qv);

      seq[rlen+1] = seq[rlen+2] = 0;
      if (arrow) orc_number_arrow(seq); else orc_number_read(seq);
      orc_compress_read((int) rlen,seq);
      put(&o,seq,(rlen+3)>>2);
    }
done:
  free(seq);
  free(hdr);
  if (o.n < 0) return o.n;
  return o.over ? ORC_E_CAP : o.n;
}

/* ------------------------------------------------------------------------------------------
 *  undexta / undexar  (undexta.c:130-271, undexar.c:130-229)
 * ---------------------------------------------------------------------------------------- */

static u16 swap16(u16 v) { return (u16) ((v >> 8) | (v << 8)); }
static u32 swap32(u32 v)
{ return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

typedef struct { const u8 *p; i64 n, at; int bad; } ibuf;

static void get(ibuf *b, void *dst, i64 k)
{ if (b->at + k > b->n)
    { b->bad = 1; memset(dst,0,(size_t) k); return; }
  memcpy(dst,b->p+b->at,(size_t) k);
  b->at += k;
}

/* well-delta, then beg/end/(qv): 32-bit fields in the new format, 16-bit in the old */
static int get_coords(ibuf *b, int newv, int flip, int *well, int *beg, int *end, int *qv)
{ u8 byte;

  if (b->at >= b->n) return 0;
  get(b,&byte,1);
  while (byte == 255)
    { *well += 255;
      get(b,&byte,1);
      if (b->bad) return -1;
    }
  *well += byte;
  if (newv)
    { u32 v[3];
      int k, m = (qv == NULL ? 2 : 3);
      get(b,v,4*m);
      for (k = 0; k < m; k++)
        if (flip) v[k] = swap32(v[k]);
      *beg = (int) v[0]; *end = (int) v[1];
      if (qv != NULL) *qv = (int) v[2];
    }
  else
    { u16 v[3];
      int k;
      get(b,v,6);
      for (k = 0; k < 3; k++)
        if (flip) v[k] = swap16(v[k]);
      *beg = v[0]; *end = v[1]; *qv = v[2];
    }
  return b->bad ? -1 : 1;
}

static void put_str(obuf *o, const char *s) { put(o,s,(i64) strlen(s)); }

i64 orc_undexta(const u8 *in, i64 n, int arrow, int width, int upper, u8 *out, i64 cap)
{ obuf  o = { out, 0, cap, 0 };
  ibuf  b = { in, n, 0, 0 };
  u16   key;
  int   flip, newv, plen, well = 0, r;
  char *name, *read, line[256];
  i64   rmax = 1<<16;

  get(&b,&key,2);
  if (b.bad) return ORC_E_TRUNC;
  if      (key == 0x55aa)           { flip = 0; newv = 1; }
  else if (key == 0xaa55)           { flip = 1; newv = 1; }
  else if (!arrow && key == 0x33cc) { flip = 0; newv = 0; }   /* undexta.c:140-155 only */
  else if (!arrow && key == 0xcc33) { flip = 1; newv = 0; }
  else return ORC_E_KEY;

  get(&b,&plen,4);
  if (flip) plen = (int) swap32((u32) plen);
  if (b.bad || plen < 0 || plen > n) return ORC_E_TRUNC;
  name = malloc((size_t) plen+1);
  get(&b,name,plen);
  name[plen] = '\0';
  if (b.bad) { free(name); return ORC_E_TRUNC; }

  read = malloc((size_t) rmax+8);
  while (1)
    { int beg, end, qv = 0, rlen, clen, j;
      u16 cnr[4];

      r = get_coords(&b,newv,flip,&well,&beg,&end,arrow ? NULL : &qv);
      if (r == 0) break;
      if (r < 0) { o.n = ORC_E_TRUNC; break; }
      if (arrow)
        { float snr[4];
          get(&b,cnr,8);
          if (b.bad) { o.n = ORC_E_TRUNC; break; }
          for (j = 0; j < 4; j++)
            { if (flip) cnr[j] = swap16(cnr[j]);
              snr[j] = cnr[j]/100.;                    /* undexar.c:199-200 (float) */
            }
          snprintf(line,sizeof(line),"/%d/%d_%d SN=%.2f,%.2f,%.2f,%.2f\n",well,beg,end,
                                     snr[0],snr[1],snr[2],snr[3]);
        }
      else
        snprintf(line,sizeof(line),"/%d/%d_%d RQ=0.%d\n",well,beg,end,qv);
      put_str(&o,name);
      put_str(&o,line);

      rlen = end-beg;
      if (rlen < 0) { o.n = ORC_E_FORMAT; break; }
      if (rlen+8 > rmax)
        { rmax = 2*(i64) rlen + 8;
          read = realloc(read,(size_t) rmax+8);
        }
      clen = (rlen+3) >> 2;
      get(&b,read,clen);
      if (b.bad) { o.n = ORC_E_TRUNC; break; }
      orc_uncompress_read(rlen,read);
      if (arrow)      orc_letter_arrow(read);
      else if (upper) orc_upper_read(read);
      else            orc_lower_read(read);
      for (j = 0; j < rlen; j += width)                /* undexta.c:263-270 */
        { int w = (j+width > rlen ? rlen-j : width);
          put(&o,read+j,w);
          put_u8(&o,'\n');
        }
    }
  free(read);
  free(name);
  if (o.n < 0) return o.n;
  return o.over ? ORC_E_CAP : o.n;
}

/* ------------------------------------------------------------------------------------------
 *  Huffman scheme construction  (QV.c:91-220)
 * ---------------------------------------------------------------------------------------- */

typedef struct { int lft, rgt, sym; u64 count; } hnode;     /* rgt < 0 marks a leaf */

static void sift(int s, int *heap, int hsize, const hnode *nd)       /* QV.c:91-120 */
{ int c = s, hs = heap[s], l;

  while ((l = 2*c) <= hsize)
    { int r = l+1, pick;
      /* left child unless the right one exists and is not larger: ties go right */
      if (r > hsize || nd[heap[r]].count > nd[heap[l]].count)
        pick = l;
      else
        pick = r;
      if (nd[hs].count > nd[heap[pick]].count)
        { heap[c] = heap[pick];
          c = pick;
        }
      else
        break;
    }
  heap[c] = hs;
}

static void assign(const hnode *nd, int v, u32 code, int len, orc_scheme *s)   /* QV.c:125-137 */
{ if (nd[v].rgt < 0)
    { s->bits[nd[v].sym] = code;
      s->lens[nd[v].sym] = len;
    }
  else
    { assign(nd,nd[v].lft,code<<1,len+1,s);
      assign(nd,nd[v].rgt,(code<<1)+1,len+1,s);
    }
}

void orc_huffman(const u64 *hist, const orc_scheme *in, orc_scheme *out)      /* QV.c:147-220 */
{ hnode nd[520];
  int   heap[260];
  int   hsize = 0, nleaf = 0, top, i;

  if (in != NULL)                      /* escape leaf goes in first */
    { nd[0].count = 0; nd[0].sym = 255; nd[0].lft = nd[0].rgt = -1;
      heap[++hsize] = nleaf++;
    }
  for (i = 0; i < 256; i++)
    if (hist[i] > 0)
      { if (in != NULL && (in->lens[i] > CUTOFF || i == 255))
          nd[0].count += hist[i];
        else
          { nd[nleaf].count = hist[i]; nd[nleaf].sym = i;
            nd[nleaf].lft = nd[nleaf].rgt = -1;
            heap[++hsize] = nleaf++;
          }
      }

  for (i = hsize/2; i >= 1; i--)
    sift(i,heap,hsize,nd);

  top = nleaf;
  for (i = 1; i < nleaf; i++)
    { int lft = heap[1], rgt;
      heap[1] = heap[hsize--];
      sift(1,heap,hsize,nd);
      rgt = heap[1];
      nd[top].lft = lft; nd[top].rgt = rgt; nd[top].sym = -1;
      nd[top].count = nd[lft].count + nd[rgt].count;
      heap[1] = top++;
      sift(1,heap,hsize,nd);
    }

  memset(out->bits,0,sizeof(out->bits));
  memset(out->lens,0,sizeof(out->lens));
  if (top > 0)
    assign(nd,top-1,0,0,out);

  if (in != NULL)
    { out->type = 2;
      for (i = 0; i < 255; i++)
        if (in->lens[i] > CUTOFF || out->lens[i] > CUTOFF)
          { out->lens[i] = out->lens[255];
            out->bits[i] = out->bits[255];
          }
    }
  else
    { out->type = 0;
      for (i = 0; i < 256; i++)
        if (out->lens[i] > CUTOFF)
          out->type = 1;
    }
}

static void make_scheme(const u64 *hist, orc_scheme *out)        /* QV.c:1069-1078 */
{ orc_scheme first;

  orc_huffman(hist,NULL,&first);
  if (first.type)
    orc_huffman(hist,&first,out);
  else
    *out = first;
}

/* ------------------------------------------------------------------------------------------
 *  .quiva text walking and the statistics scan  (QV.c:751-798, 922-1023)
 * ---------------------------------------------------------------------------------------- */

typedef struct { i64 hdr, hlen, line[5]; int rlen; int well, beg, end, qv; } qentry;

/* next entry at *pos.  1 ok, 0 clean end of input, <0 error */
static int next_qentry(const u8 *t, i64 n, i64 *pos, qentry *e, int need_fields)
{ i64  p = *pos, q;
  int  k;
  char hdr[1024];

  if (p >= n) return 0;
  for (q = p; q < n && t[q] != '\n'; q++)
    ;
  if (q >= n) return ORC_E_FORMAT;                     /* QV.c:778-781 */
  e->hdr = p; e->hlen = q-p;
  if (e->hlen == 0 || t[p] != '@') return ORC_E_FORMAT; /* QV.c:954 */
  if (need_fields)
    { i64   m = (e->hlen < 1000 ? e->hlen : 1000);
      char *slash;
      memcpy(hdr,t+p,(size_t) m); hdr[m] = '\n'; hdr[m+1] = '\0';
      slash = strchr(hdr+1,'/');
      if (slash == NULL) return ORC_E_FORMAT;
      if (sscanf(slash+1,"%d/%d_%d RQ=0.%d\n",&e->well,&e->beg,&e->end,&e->qv) != 4)
        return ORC_E_FORMAT;
    }
  p = q+1;
  for (k = 0; k < 5; k++)
    { if (p >= n) return ORC_E_FORMAT;                 /* incomplete last entry */
      for (q = p; q < n && t[q] != '\n'; q++)
        ;
      if (q >= n) return ORC_E_FORMAT;
      if (k == 0) e->rlen = (int) (q-p);
      else if (q-p != e->rlen) return ORC_E_LINELEN;   /* QV.c:792 */
      e->line[k] = p;
      p = q+1;
    }
  *pos = p;
  return 1;
}

static void count_syms(u64 *h, const u8 *s, int len)             /* QV.c:702-707 */
{ int k;
  for (k = 0; k < len; k++) h[s[k]] += 1;
}

static void count_runs(u64 *run, const u8 *s, int len, int rc)   /* QV.c:709-724 */
{ int k = 0;
  while (k < len)
    { int h = k;
      while (k < len && s[k] == rc) k++;
      run[k-h >= 255 ? 255 : k-h] += 1;
      if (k < len) k++;
    }
}

int orc_qv_scan(const u8 *t, i64 n, orc_stats *st)               /* QV.c:922-1023 */
{ i64    pos = 0;
  qentry e;
  int    r, k;

  memset(st,0,sizeof(*st));
  for (k = 0; k < 256; k++) st->delrun[k] = st->subrun[k] = 1;
  st->delchar = st->subchar = -1;

  while ((r = next_qentry(t,n,&pos,&e,1)) > 0)
    { const u8 *del = t+e.line[0], *tag = t+e.line[1];
      const u8 *sub = t+e.line[4];

      count_syms(st->del,del,e.rlen);
      count_syms(st->ins,t+e.line[2],e.rlen);
      count_syms(st->mrg,t+e.line[3],e.rlen);
      count_syms(st->sub,sub,e.rlen);
      if (st->delchar < 0)
        for (k = 0; k < e.rlen; k++)
          if (tag[k] == 'n' || tag[k] == 'N')
            { st->delchar = del[k]; break; }
      if (st->delchar >= 0)
        count_runs(st->delrun,del,e.rlen,st->delchar);
      st->totchar += (u64) e.rlen;
      if (st->subchar < 0 && st->totchar >= 100000)
        { st->subchar = 0;
          for (k = 1; k < 256; k++)
            if (st->sub[k] > st->sub[st->subchar]) st->subchar = k;
        }
      if (st->subchar >= 0)
        count_runs(st->subrun,sub,e.rlen,st->subchar);
      st->nentries += 1;
    }
  return r;
}

int orc_qv_create(orc_stats *st, int lossy, orc_coding *c)       /* QV.c:1029-1169 */
{ int k;

  if (st->totchar < 200000 || st->sub[st->subchar < 0 ? 0 : st->subchar] < .5*st->totchar)
    st->subchar = -1;
  if (lossy)
    { for (k = 0; k < 256; k += 2)
        { st->ins[k] += st->ins[k+1]; st->ins[k+1] = 0; }
      for (k = 0; k < 256; k += 4)
        { st->mrg[k] += st->mrg[k+1] + st->mrg[k+2] + st->mrg[k+3];
          st->mrg[k+1] = st->mrg[k+2] = st->mrg[k+3] = 0;
        }
    }
  memset(c,0,sizeof(*c));
  if (st->delchar >= 0)
    { st->del[st->delchar] = 0;
      make_scheme(st->delrun,&c->tab[1]);
    }
  make_scheme(st->del,&c->tab[0]);
  make_scheme(st->ins,&c->tab[2]);
  make_scheme(st->mrg,&c->tab[3]);
  if (st->subchar >= 0)
    { st->sub[st->subchar] = 0;
      make_scheme(st->subrun,&c->tab[5]);
    }
  make_scheme(st->sub,&c->tab[4]);
  c->delchar = st->delchar;
  c->subchar = st->subchar;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 *  Coding header (QV.c:300-318, 1173-1210; reader 322-375, 1214-1320)
 * ---------------------------------------------------------------------------------------- */

static void put_scheme(obuf *o, const orc_scheme *s)
{ int i;
  put_u8(o,(u8) s->type);
  for (i = 0; i < 256; i++)
    { put_u8(o,(u8) s->lens[i]);
      if ((u8) s->lens[i] > 0)
        put(o,&s->bits[i],4);
    }
}

i64 orc_write_coding(const orc_coding *c, const char *prefix, int plen, u8 *out, i64 cap)
{ obuf o = { out, 0, cap, 0 };

  put_u16(&o,0x33cc);
  put_u16(&o,(u16) (c->delchar < 0 ? 256 : c->delchar));
  put_u16(&o,(u16) (c->subchar < 0 ? 256 : c->subchar));
  put_i32(&o,plen);
  put(&o,prefix,plen);
  put_scheme(&o,&c->tab[0]);
  if (c->delchar >= 0) put_scheme(&o,&c->tab[1]);
  put_scheme(&o,&c->tab[2]);
  put_scheme(&o,&c->tab[3]);
  put_scheme(&o,&c->tab[4]);
  if (c->subchar >= 0) put_scheme(&o,&c->tab[5]);
  return o.over ? ORC_E_CAP : o.n;
}

static void get_scheme(ibuf *b, orc_scheme *s, int flip)
{ int i;
  u8  x;
  get(b,&x,1);
  s->type = x;
  for (i = 0; i < 256; i++)
    { get(b,&x,1);
      s->lens[i] = x;
      s->bits[i] = 0;
      if (x > 0)
        { get(b,&s->bits[i],4);
          if (flip) s->bits[i] = swap32(s->bits[i]);
        }
    }
}

i64 orc_read_coding(const u8 *in, i64 n, orc_coding *c, char *prefix, int pcap, int *flip)
{ ibuf b = { in, n, 0, 0 };
  u16  half;
  int  len;

  memset(c,0,sizeof(*c));
  get(&b,&half,2);
  *flip = (half != 0x33cc);
  get(&b,&half,2); if (*flip) half = swap16(half);
  c->delchar = (half >= 256 ? -1 : half);
  get(&b,&half,2); if (*flip) half = swap16(half);
  c->subchar = (half >= 256 ? -1 : half);
  get(&b,&len,4);  if (*flip) len = (int) swap32((u32) len);
  if (b.bad || len < 0 || len >= pcap) return ORC_E_TRUNC;
  get(&b,prefix,len);
  prefix[len] = '\0';
  get_scheme(&b,&c->tab[0],*flip);
  if (c->delchar >= 0) get_scheme(&b,&c->tab[1],*flip);
  get_scheme(&b,&c->tab[2],*flip);
  get_scheme(&b,&c->tab[3],*flip);
  get_scheme(&b,&c->tab[4],*flip);
  if (c->subchar >= 0) get_scheme(&b,&c->tab[5],*flip);
  return b.bad ? ORC_E_TRUNC : b.at;
}

/* ------------------------------------------------------------------------------------------
 *  Bit-stream encoder  (QV.c:386-506), stated as "items into a zeroed word array"
 * ---------------------------------------------------------------------------------------- */

typedef struct { u32 *w; i64 bits, last; int any; } bitw;

static void item(bitw *b, int len, u32 code)
{ i64 word = b->bits >> 5;
  int off  = (int) (b->bits & 31);

  b->last = b->bits;
  b->any  = 1;
  if (len <= 0) return;            /* absent symbol: not a parity target, just stay defined */
  code &= (len >= 32 ? 0xffffffffu : ((1u << len)-1));
  if (off + len <= 32)
    b->w[word] |= code << (32-off-len);
  else
    { b->w[word]   |= code >> (off+len-32);
      b->w[word+1] |= code << (64-off-len);
    }
  b->bits += len;
}

/* number of 32-bit words the reference writes for the stream, filling the possible extra
 * word (QV.c:436-442 / 499-505): a copy of the trailing partial word, or 0 on a word boundary */
static i64 finish(bitw *b)
{ i64 full, nw;

  if (!b->any) return 0;
  full = (b->bits+31) >> 5;
  nw   = (b->last+47) >> 5;
  if (nw < full) nw = full;
  if (nw > full)
    b->w[full] = (b->bits & 31) ? b->w[full-1] : 0;
  return nw;
}

i64 orc_encode_stream(const orc_scheme *sym, const orc_scheme *run, int rchar,
                      const u8 *s, int rlen, u8 *out, i64 cap)
{ bitw b;
  i64  nw, maxw = ((i64) rlen*40 + 64)/32 + 4;
  int  k;
  u32  nspec = 0x7fffffff, rspec = 0;
  int  nslen = 0x7fffffff, rslen = 0;

  b.w = calloc((size_t) maxw,4);
  b.bits = b.last = 0; b.any = 0;
  if (sym->type == 2) { nspec = sym->bits[255]; nslen = sym->lens[255]; }

  if (run == NULL)
    for (k = 0; k < rlen; k++)
      { int x = s[k];
        item(&b,sym->lens[x],sym->bits[x]);
        if (sym->bits[x] == nspec && sym->lens[x] == nslen)
          item(&b,8,(u32) x);
      }
  else
    { rspec = run->bits[255]; rslen = run->lens[255];
      k = 0;
      while (k < rlen)
        { int h = k, x;
          while (k < rlen && s[k] == rchar) k++;
          x = (k-h >= 255 ? 255 : k-h);
          item(&b,run->lens[x],run->bits[x]);
          if (run->bits[x] == rspec && run->lens[x] == rslen)
            item(&b,16,(u32) (k-h));
          if (k < rlen)
            { x = s[k];
              item(&b,sym->lens[x],sym->bits[x]);
              if (sym->bits[x] == nspec && sym->lens[x] == nslen)
                item(&b,8,(u32) x);
              k++;
            }
        }
    }
  nw = finish(&b);
  if (nw*4 > cap) { free(b.w); return ORC_E_CAP; }
  memcpy(out,b.w,(size_t) nw*4);
  free(b.w);
  return nw*4;
}

/* ------------------------------------------------------------------------------------------
 *  dexqv  (dexqv.c:59-147 driving QV.c:1381-1426)
 * ---------------------------------------------------------------------------------------- */

i64 orc_dexqv(const u8 *t, i64 n, int lossy, u8 *out, i64 cap)
{ orc_stats  *st = malloc(sizeof(orc_stats));
  orc_coding *c  = malloc(sizeof(orc_coding));
  obuf   o = { out, 0, cap, 0 };
  qentry e;
  i64    pos, r;
  int    lwell = 0, k;
  u8    *tmp = NULL, *line = NULL;
  i64    tmax = 0;

  r = orc_qv_scan(t,n,st);
  if (r < 0 || st->nentries == 0) { free(st); free(c); return (r < 0 ? r : ORC_E_FORMAT); }
  orc_qv_create(st,lossy,c);

  { const u8 *slash = memchr(t+1,'/',(size_t) (n-1));           /* dexqv.c:88-103 */
    i64 w;
    put_u16(&o,0x55aa);
    w = orc_write_coding(c,(const char *) t,(int) (slash-t),out+o.n,cap-o.n);
    if (w < 0) { free(st); free(c); return w; }
    o.n += w;
  }

  pos = 0;
  while ((r = next_qentry(t,n,&pos,&e,1)) > 0)
    { int rlen = e.rlen, clen;
      i64 need = (i64) rlen*5 + 64, w;

      if (need > tmax)
        { tmax = 2*need;
          tmp  = realloc(tmp,(size_t) tmax);
          line = realloc(line,(size_t) rlen*2+16);
        }
      put_well(&o,e.well,&lwell);
      put_i32(&o,e.beg); put_i32(&o,e.end); put_i32(&o,e.qv);

      /* deletion QVs, then the tags that survive (QV.c:1393-1404) */
      if (c->delchar < 0)
        w = orc_encode_stream(&c->tab[0],NULL,-1,t+e.line[0],rlen,tmp,tmax);
      else
        w = orc_encode_stream(&c->tab[0],&c->tab[1],c->delchar,t+e.line[0],rlen,tmp,tmax);
      put(&o,tmp,w);
      clen = 0;
      for (k = 0; k < rlen; k++)
        if (c->delchar < 0 || t[e.line[0]+k] != c->delchar)     /* QV.c:810-819 */
          line[clen++] = t[e.line[1]+k];
      line[clen] = '\0'; line[clen+1] = line[clen+2] = 0;
      orc_number_read((char *) line);
      orc_compress_read(clen,(char *) line);
      put(&o,line,(clen+3)>>2);

      for (k = 2; k <= 3; k++)                                  /* QV.c:1406-1418 */
        { const u8 *src = t+e.line[k];
          int j;
          if (lossy)
            { for (j = 0; j < rlen; j++)
                line[j] = (k == 2) ? (u8) ((src[j] >> 1) << 1) : (u8) ((src[j] >> 2) << 2);
              src = line;
            }
          w = orc_encode_stream(&c->tab[k],NULL,-1,src,rlen,tmp,tmax);
          put(&o,tmp,w);
        }
      if (c->subchar < 0)                                       /* QV.c:1419-1423 */
        w = orc_encode_stream(&c->tab[4],NULL,-1,t+e.line[4],rlen,tmp,tmax);
      else
        w = orc_encode_stream(&c->tab[4],&c->tab[5],c->subchar,t+e.line[4],rlen,tmp,tmax);
      put(&o,tmp,w);
    }
  free(tmp); free(line); free(st); free(c);
  if (r < 0) return r;
  return o.over ? ORC_E_CAP : o.n;
}

/* ------------------------------------------------------------------------------------------
 *  Bit-stream decoder  (QV.c:510-691) and undexqv (undexqv.c:99-208, QV.c:1428-1481)
 * ---------------------------------------------------------------------------------------- */

typedef struct { u8 look[6][65536]; } lut_t;

static void build_lut(const orc_scheme *s, u8 *look)             /* QV.c:365-372 */
{ int i;
  u32 j;
  memset(look,0,65536);
  for (i = 0; i < 256; i++)
    if (s->lens[i] > 0 && s->lens[i] <= 16)
      { u32 base = (s->bits[i] << (16-s->lens[i])) & 0xffff;
        u32 span = 1u << (16-s->lens[i]);
        for (j = 0; j < span; j++) look[base+j] = (u8) i;
      }
}

/* A 16-bit look-ahead window over whole 32-bit words: a new word is fetched exactly when the
 * window would run past the words fetched so far (the GET macro, QV.c:537-551). */
typedef struct { ibuf *b; int flip; u64 acc; i64 pos, loaded; } bitr;

static void advance(bitr *r, int nbits)
{ r->pos += nbits;
  if (r->pos + 16 > r->loaded)
    { u32 w;
      get(r->b,&w,4);
      if (r->flip) w = swap32(w);
      r->acc = (r->acc << 32) | w;
      r->loaded += 32;
    }
}

static u32 window(const bitr *r)        /* the 16 bits at pos */
{ int sh = (int) (r->loaded - r->pos - 16);
  return (u32) ((r->acc >> sh) & 0xffff);
}

static int decode_stream(ibuf *b, int flip, const orc_scheme *sym, const u8 *slook,
                         const orc_scheme *run, const u8 *rlook, int rchar, u8 *dst, int rlen)
{ bitr r = { b, flip, 0, -16, 0 };
  int  signal = (sym->type == 2 ? 255 : 256);
  int  n = 16, j, c, k;

  if (run == NULL)
    for (j = 0; j < rlen; j++)
      { advance(&r,n);
        c = slook[window(&r)];
        n = sym->lens[c];
        if (c == signal)
          { advance(&r,n);
            c = (int) (window(&r) >> 8);
            n = 8;
          }
        dst[j] = (u8) c;
      }
  else
    for (j = 0; j < rlen; j++)
      { advance(&r,n);
        c = rlook[window(&r)];
        n = run->lens[c];
        if (c == 255)
          { advance(&r,n);
            c = (int) window(&r);
            n = 16;
          }
        for (k = 0; k < c && j < rlen; k++)
          dst[j++] = (u8) rchar;
        if (k < c) return ORC_E_FORMAT;
        if (j < rlen)
          { advance(&r,n);
            c = slook[window(&r)];
            n = sym->lens[c];
            if (c == signal)
              { advance(&r,n);
                c = (int) (window(&r) >> 8);
                n = 8;
              }
            dst[j] = (u8) c;
          }
      }
  return b->bad ? ORC_E_TRUNC : 0;
}

/* shared by orc_undexqv and orc_dexqv_offsets */
static i64 walk_dexqv(const u8 *in, i64 n, int upper, obuf *o, i64 *offs, i64 maxent)
{ ibuf  b = { in, n, 0, 0 };
  orc_coding *c = malloc(sizeof(orc_coding));
  lut_t *lut = malloc(sizeof(lut_t));
  char  prefix[4096], line[256];
  int   flip, newv, well = 0, k, r;
  u16   key;
  i64   w, nent = 0, ret = 0;
  u8   *ent = NULL;
  i64   emax = 0;

  get(&b,&key,2);
  if (b.bad) { ret = ORC_E_TRUNC; goto out; }
  if (key == 0x55aa || key == 0xaa55) newv = 1;         /* undexqv.c:103-110 */
  else { newv = 0; b.at = 0; }
  w = orc_read_coding(in+b.at,n-b.at,c,prefix,sizeof(prefix),&flip);
  if (w < 0) { ret = w; goto out; }
  b.at += w;
  for (k = 0; k < 6; k++) build_lut(&c->tab[k],lut->look[k]);

  while (1)
    { int beg, end, qv, rlen, clen, tlen;
      i64 at0 = b.at;

      r = get_coords(&b,newv,flip,&well,&beg,&end,&qv);
      if (r == 0) break;
      if (r < 0) { ret = ORC_E_TRUNC; break; }
      if (offs != NULL)
        { if (nent >= maxent) { ret = ORC_E_CAP; break; }
          offs[nent] = at0;
        }
      nent += 1;
      rlen = end-beg;
      if (rlen < 0) { ret = ORC_E_FORMAT; break; }
      if ((i64) rlen+8 > emax)
        { emax = 2*(i64) rlen + 64;
          ent  = realloc(ent,(size_t) (5*emax));
        }
      if (o != NULL)
        { snprintf(line,sizeof(line),"/%d/%d_%d RQ=0.%d\n",well,beg,end,qv);
          put_str(o,prefix);
          put_str(o,line);
        }

      if (c->delchar < 0)                                /* QV.c:1433-1462 */
        { r = decode_stream(&b,flip,&c->tab[0],lut->look[0],NULL,NULL,-1,ent,rlen);
          clen = rlen;
        }
      else
        { r = decode_stream(&b,flip,&c->tab[0],lut->look[0],&c->tab[1],lut->look[1],
                            c->delchar,ent,rlen);
          clen = 0;
          for (k = 0; k < rlen; k++) clen += (ent[k] != c->delchar);
        }
      if (r < 0) { ret = r; break; }
      tlen = (clen+3) >> 2;
      { u8 *tag = ent+emax;
        get(&b,tag,tlen);
        if (b.bad) { ret = ORC_E_TRUNC; break; }
        orc_uncompress_read(clen,(char *) tag);
        orc_lower_read((char *) tag);
        if (c->delchar >= 0)                             /* QV.c:837-847 */
          { int j = clen-1;
            for (k = rlen-1; k >= 0; k--)
              tag[k] = (ent[k] == c->delchar) ? 'n' : tag[j--];
          }
        if (upper)
          for (k = 0; k < rlen; k++) tag[k] -= 32;       /* undexqv.c:198-204 */
      }
      r = decode_stream(&b,flip,&c->tab[2],lut->look[2],NULL,NULL,-1,ent+2*emax,rlen);
      if (r < 0) { ret = r; break; }
      r = decode_stream(&b,flip,&c->tab[3],lut->look[3],NULL,NULL,-1,ent+3*emax,rlen);
      if (r < 0) { ret = r; break; }
      if (c->subchar < 0)
        r = decode_stream(&b,flip,&c->tab[4],lut->look[4],NULL,NULL,-1,ent+4*emax,rlen);
      else
        r = decode_stream(&b,flip,&c->tab[4],lut->look[4],&c->tab[5],lut->look[5],
                          c->subchar,ent+4*emax,rlen);
      if (r < 0) { ret = r; break; }
      if (o != NULL)
        for (k = 0; k < 5; k++)
          { put(o,ent+k*emax,rlen);
            put_u8(o,'\n');
          }
    }
  if (ret == 0 && offs != NULL)
    { if (nent >= maxent) ret = ORC_E_CAP; else offs[nent] = b.at; }
out:
  free(ent); free(lut); free(c);
  return ret < 0 ? ret : nent;
}

i64 orc_undexqv(const u8 *in, i64 n, int upper, u8 *out, i64 cap)
{ obuf o = { out, 0, cap, 0 };
  i64  r = walk_dexqv(in,n,upper,&o,NULL,0);
  if (r < 0) return r;
  return o.over ? ORC_E_CAP : o.n;
}

i64 orc_dexqv_offsets(const u8 *in, i64 n, i64 *offs, i64 maxent)
{ return walk_dexqv(in,n,0,NULL,offs,maxent+1); }
