/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
 *
 * dx_oracle: a sequential plain-C restatement, buffer-to-buffer, of the DEXTRACTOR compression
 * hot path (2-bit packing behind dexta/dexar, Huffman + run-length QV coder behind dexqv).
 * Every function cites the reference lines it restates (paths are into /root/reference).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py runs this library against the reference
 * tools themselves (oracle/_ref, compiled by oracle/Makefile from the mounted sources) on seeded
 * inputs, and tests/test_golden.py checks it against committed reference outputs in
 * tests/golden/ (made by tests/golden/make_golden.py with those same reference binaries).
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py
 * may load this library.
 */
#ifndef DX_ORACLE_H
#define DX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (all negative) */
#define ORC_E_FORMAT   (-1)   /* malformed text / header line                       */
#define ORC_E_CAP      (-2)   /* output buffer too small                            */
#define ORC_E_TRUNC    (-3)   /* compressed input ends early                        */
#define ORC_E_KEY      (-4)   /* bad endian key                                     */
#define ORC_E_LINELEN  (-5)   /* quiva lines of one entry differ in length          */
#define ORC_E_TOOLONG  (-6)   /* fasta/arrow line longer than the reference allows  */

/* Statistics gathered by the scan (reference QV.c:860-862 static state). */
typedef struct
  { uint64_t del[256], ins[256], mrg[256], sub[256], delrun[256], subrun[256];
    uint64_t totchar;
    int32_t  delchar, subchar;     /* -1 when undetermined */
    int64_t  nentries;
  } orc_stats;

/* One Huffman table (reference QV.c:76-81 HScheme, without the decode LUT). */
typedef struct
  { int32_t  type;
    uint32_t bits[256];
    int32_t  lens[256];
  } orc_scheme;

/* reference QV.h:31-42 QVcoding, with concrete tables. index: 0 del 1 drun 2 ins 3 mrg 4 sub 5 srun */
typedef struct
  { orc_scheme tab[6];
    int32_t    delchar, subchar;
  } orc_coding;

/* ---- 2-bit codec primitives (reference DB.c:319-441) ---- */
void orc_number_read(char *s);                 /* DB.c:393-416 */
void orc_number_arrow(char *s);                /* DB.c:418-441 */
void orc_compress_read(int len, char *s);      /* DB.c:319-338 */
void orc_uncompress_read(int len, char *s);    /* DB.c:342-363 */
void orc_lower_read(char *s);                  /* DB.c:367-373 */
void orc_upper_read(char *s);                  /* DB.c:375-381 */
void orc_letter_arrow(char *s);                /* DB.c:383-389 */

/* ---- whole-file tools, buffer to buffer; return output length or a negative error ---- */
int64_t orc_dexta  (const uint8_t *text, int64_t n, int arrow, uint8_t *out, int64_t cap);
int64_t orc_undexta(const uint8_t *in, int64_t n, int arrow, int width, int upper,
                    uint8_t *out, int64_t cap);
int64_t orc_dexqv  (const uint8_t *text, int64_t n, int lossy, uint8_t *out, int64_t cap);
int64_t orc_undexqv(const uint8_t *in, int64_t n, int upper, uint8_t *out, int64_t cap);

/* ---- pieces of the QV coder, exposed for unit tests ---- */
int     orc_qv_scan(const uint8_t *text, int64_t n, orc_stats *st);          /* QV.c:922-1023 */
int     orc_qv_create(orc_stats *st, int lossy, orc_coding *c);              /* QV.c:1029-1169 */
void    orc_huffman(const uint64_t *hist, const orc_scheme *in, orc_scheme *out); /* QV.c:147-220 */
int64_t orc_write_coding(const orc_coding *c, const char *prefix, int plen,
                         uint8_t *out, int64_t cap);                          /* QV.c:1173-1210 */
int64_t orc_read_coding(const uint8_t *in, int64_t n, orc_coding *c,
                        char *prefix, int pcap, int *flip);                   /* QV.c:1214-1320 */
/* encode one stream; run<0 => plain (QV.c:386-443), else run coding (QV.c:448-506) */
int64_t orc_encode_stream(const orc_scheme *sym, const orc_scheme *run, int rchar,
                          const uint8_t *s, int rlen, uint8_t *out, int64_t cap);
/* per-entry byte offsets of a .dexqv (start of each entry's well-delta bytes) + end of file;
   offs must hold nentries+1 values. returns nentries or negative error. */
int64_t orc_dexqv_offsets(const uint8_t *in, int64_t n, int64_t *offs, int64_t maxent);

#ifdef __cplusplus
}
#endif
#endif
