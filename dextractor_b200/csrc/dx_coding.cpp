// dx_coding.cpp -- host side of the QV coder: code-length assignment and the coding header.
//
// The code tables must be byte-identical to the reference's, so the Huffman construction keeps
// the reference's observable choices (reference QV.c:91-220):
//   * leaves enter the heap in ascending symbol order, an escape leaf (symbol 255) first when a
//     truncated table is built;
//   * sift-down prefers the RIGHT child on equal counts and only moves on a strict '>' ;
//   * a merge pops the minimum, then REPLACES the new root with the merged node;
//   * codes: left edge 0, right edge 1, most significant bit first.
// Everything here runs in microseconds on the host; the GPU never sees a tree, only the tables.

#include <string.h>
#include <stdlib.h>
#include "dx_internal.h"

namespace {

constexpr int kCutoff = 16;       // QV.c:26: codes longer than this fold into the escape

struct Forest
{ uint64_t weight[520];
  int16_t  kid0[520], kid1[520];   // kid1 < 0: leaf, kid0 is then the symbol
  int      heap[260];
  int      nheap = 0, nnode = 0;

  int leaf(int symbol, uint64_t w)
  { weight[nnode] = w; kid0[nnode] = (int16_t) symbol; kid1[nnode] = -1;
    heap[++nheap] = nnode;
    return nnode++;
  }

  void settle(int at)             // QV.c:91-120
  { const int moving = heap[at];
    int hole = at;
    for (int l = 2*hole; l <= nheap; l = 2*hole)
      { const int r = l+1;
        const int pick = (r > nheap || weight[heap[r]] > weight[heap[l]]) ? l : r;
        if (!(weight[moving] > weight[heap[pick]]))
          break;
        heap[hole] = heap[pick];
        hole = pick;
      }
    heap[hole] = moving;
  }

  int build()                     // QV.c:180-194; returns the root node, -1 if no leaves
  { const int nleaf = nnode;
    for (int i = nheap/2; i >= 1; i--)
      settle(i);
    for (int i = 1; i < nleaf; i++)
      { const int a = heap[1];
        heap[1] = heap[nheap--];
        settle(1);
        const int b = heap[1];
        weight[nnode] = weight[a] + weight[b];
        kid0[nnode] = (int16_t) a; kid1[nnode] = (int16_t) b;
        heap[1] = nnode++;
        settle(1);
      }
    return nnode-1;
  }

  void label(int node, uint32_t code, int depth, dx_scheme *s) const   // QV.c:125-137
  { if (kid1[node] < 0)
      { s->bits[kid0[node]] = code;
        s->lens[kid0[node]] = depth;
        return;
      }
    label(kid0[node], code << 1,       depth+1, s);
    label(kid1[node], (code << 1) | 1, depth+1, s);
  }
};

// One pass of the construction (QV.c:147-220).  `first` is the untruncated table of the same
// histogram when the truncated (type 2) variant is wanted, else NULL.
void construct(const uint64_t *hist, const dx_scheme *first, dx_scheme *out)
{ Forest f;
  int    esc = -1;

  if (first != NULL)
    esc = f.leaf(255,0);
  for (int i = 0; i < 256; i++)
    { if (hist[i] == 0) continue;
      if (first != NULL && (first->lens[i] > kCutoff || i == 255))
        f.weight[esc] += hist[i];
      else
        f.leaf(i,hist[i]);
    }
  memset(out,0,sizeof(*out));
  const int root = f.build();
  if (root >= 0)
    f.label(root,0,0,out);

  if (first != NULL)
    { out->type = 2;
      for (int i = 0; i < 255; i++)
        if (first->lens[i] > kCutoff || out->lens[i] > kCutoff)
          { out->lens[i] = out->lens[255];
            out->bits[i] = out->bits[255];
          }
    }
  else
    { out->type = 0;
      for (int i = 0; i < 256; i++)
        if (out->lens[i] > kCutoff)
          out->type = 1;
    }
}

int table_for(const uint64_t *hist, dx_scheme *out)       // SCHEME_MACRO, QV.c:1069-1078
{ int distinct = 0;
  for (int i = 0; i < 256; i++)
    distinct += (hist[i] > 0);
  if (distinct < 2)              // the reference itself misbehaves here
    return DX_E_CODING;
  dx_scheme plain;
  construct(hist,NULL,&plain);
  if (plain.type != 0)
    construct(hist,&plain,out);
  else
    *out = plain;
  return DX_OK;
}

struct Sink
{ uint8_t *p; size_t n, cap; bool over;
  void raw(const void *src, size_t k)
  { if (n+k > cap) { over = true; return; }
    memcpy(p+n,src,k); n += k;
  }
  template <class T> void val(T v) { raw(&v,sizeof(T)); }
};

struct Source
{ const uint8_t *p; size_t n, at; bool bad;
  void raw(void *dst, size_t k)
  { if (at+k > n) { bad = true; memset(dst,0,k); return; }
    memcpy(dst,p+at,k); at += k;
  }
  template <class T> T val() { T v; raw(&v,sizeof(T)); return v; }
};

inline uint16_t bswap16(uint16_t v) { return (uint16_t) ((v >> 8) | (v << 8)); }
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

void put_table(Sink &o, const dx_scheme &s)               // Write_Scheme, QV.c:300-318
{ o.val<uint8_t>((uint8_t) s.type);
  for (int i = 0; i < 256; i++)
    { const uint8_t len = (uint8_t) s.lens[i];
      o.val<uint8_t>(len);
      if (len > 0)
        o.val<uint32_t>(s.bits[i]);
    }
}

void get_table(Source &in, dx_scheme &s, bool flip)       // Read_Scheme, QV.c:322-363
{ s.type = in.val<uint8_t>();
  for (int i = 0; i < 256; i++)
    { const uint8_t len = in.val<uint8_t>();
      s.lens[i] = len;
      s.bits[i] = 0;
      if (len > 0)
        { const uint32_t b = in.val<uint32_t>();
          s.bits[i] = flip ? bswap32(b) : b;
        }
    }
}

}  // namespace

extern "C" int dx_qv_make_coding(const dx_qv_stats *stats, int lossy, dx_qv_coding *coding)
{ if (stats == NULL || coding == NULL) return DX_E_ARG;
  // working copies: the construction edits the histograms (QV.c:1049-1065, 1102, 1129)
  uint64_t del[256], ins[256], mrg[256], sub[256], drun[256], srun[256];
  memcpy(del,stats->hist[0],sizeof(del));
  memcpy(ins,stats->hist[1],sizeof(ins));
  memcpy(mrg,stats->hist[2],sizeof(mrg));
  memcpy(sub,stats->hist[3],sizeof(sub));
  for (int i = 0; i < 256; i++)            // every run bucket starts at 1 (QV.c:934-935)
    { drun[i] = stats->hist[4][i] + 1;
      srun[i] = stats->hist[5][i] + 1;
    }

  int delchar = stats->delchar, subchar = stats->subchar;
  // a substitution run character must cover half of a reasonably large file (QV.c:1044-1045)
  if (subchar >= 0 && (stats->totchar < 200000 || (double) sub[subchar] < .5*(double) stats->totchar))
    subchar = -1;

  if (lossy)                               // QV.c:1049-1065
    { for (int k = 0; k < 256; k += 2)
        { ins[k] += ins[k+1]; ins[k+1] = 0; }
      for (int k = 0; k < 256; k += 4)
        { mrg[k] += mrg[k+1] + mrg[k+2] + mrg[k+3];
          mrg[k+1] = mrg[k+2] = mrg[k+3] = 0;
        }
    }

  memset(coding,0,sizeof(*coding));
  coding->delchar = delchar;
  coding->subchar = subchar;
  int rc;
  if (delchar >= 0)
    { del[delchar] = 0;
      if ((rc = table_for(drun,&coding->tab[1])) != DX_OK) return rc;
    }
  if ((rc = table_for(del,&coding->tab[0])) != DX_OK) return rc;
  if ((rc = table_for(ins,&coding->tab[2])) != DX_OK) return rc;
  if ((rc = table_for(mrg,&coding->tab[3])) != DX_OK) return rc;
  if (subchar >= 0)
    { sub[subchar] = 0;
      if ((rc = table_for(srun,&coding->tab[5])) != DX_OK) return rc;
    }
  if ((rc = table_for(sub,&coding->tab[4])) != DX_OK) return rc;
  return DX_OK;
}

extern "C" int dx_qv_write_coding(const dx_qv_coding *c, const char *prefix, int plen,
                                  uint8_t *out, size_t cap, size_t *out_len)
{ if (c == NULL || prefix == NULL || plen < 0 || out == NULL) return DX_E_ARG;
  Sink o = { out, 0, cap, false };
  o.val<uint16_t>(0x33cc);                                   // QV.c:1180
  o.val<uint16_t>((uint16_t) (c->delchar < 0 ? 256 : c->delchar));
  o.val<uint16_t>((uint16_t) (c->subchar < 0 ? 256 : c->subchar));
  o.val<int32_t>(plen);
  o.raw(prefix,(size_t) plen);
  put_table(o,c->tab[0]);
  if (c->delchar >= 0) put_table(o,c->tab[1]);
  put_table(o,c->tab[2]);
  put_table(o,c->tab[3]);
  put_table(o,c->tab[4]);
  if (c->subchar >= 0) put_table(o,c->tab[5]);
  if (o.over) return DX_E_CAP;
  if (out_len) *out_len = o.n;
  return DX_OK;
}

extern "C" int dx_qv_read_coding(const uint8_t *in, size_t n, dx_qv_coding *c,
                                 char *prefix, int pcap, size_t *used)
{ if (in == NULL || c == NULL || prefix == NULL || pcap < 1) return DX_E_ARG;
  Source s = { in, n, 0, false };
  memset(c,0,sizeof(*c));
  const bool flip = (s.val<uint16_t>() != 0x33cc);           // QV.c:1226
  uint16_t h = s.val<uint16_t>(); if (flip) h = bswap16(h);
  c->delchar = (h >= 256 ? -1 : h);
  h = s.val<uint16_t>(); if (flip) h = bswap16(h);
  c->subchar = (h >= 256 ? -1 : h);
  uint32_t len = s.val<uint32_t>(); if (flip) len = bswap32(len);
  if (s.bad || len >= (uint32_t) pcap || len > n) return DX_E_TRUNC;
  s.raw(prefix,len);
  prefix[len] = '\0';
  c->flip = flip;
  get_table(s,c->tab[0],flip);
  if (c->delchar >= 0) get_table(s,c->tab[1],flip);
  get_table(s,c->tab[2],flip);
  get_table(s,c->tab[3],flip);
  get_table(s,c->tab[4],flip);
  if (c->subchar >= 0) get_table(s,c->tab[5],flip);
  if (s.bad) return DX_E_TRUNC;
  if (used) *used = s.at;
  return DX_OK;
}
