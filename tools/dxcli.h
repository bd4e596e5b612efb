/* dxcli.h -- shared driver of the six command-line tools (dexta undexta dexar undexar dexqv
 * undexqv).  The tools keep the reference's command lines, flags, file naming and messages
 * (reference dexta.c:25-66, undexta.c:45-93, dexqv.c:24-53, undexqv.c:41-70 and DB.h:79-123), but
 * instead of streaming one read at a time through stdio they hand the WHOLE file to
 * libdexb200.so, which runs it on the GPU.  Host code is plain C; there is no CPU codec here. */
#ifndef DXCLI_H
#define DXCLI_H

#include <stddef.h>
#include <stdint.h>
#include "dexb200.h"

typedef struct
  { int verbose, keep, pipe, upper, lossy, width; } dx_opts;

typedef struct
  { const char *name;        /* program name for messages                       */
    const char *usage;       /* text after "Usage: <name> "                     */
    const char *flags;       /* legal single-letter flags, e.g. "vkiU"          */
    int         has_width;   /* accepts -w<int>                                 */
    const char *src_ext;     /* ".fasta" ...                                    */
    const char *dst_ext;     /* ".dexta" ...                                    */
    const char *help[5];     /* the "      -k: ..." lines of the usage message  */
    /* d_in holds n input bytes on the device; produce *out_len bytes at *d_out (device memory
       obtained with dx_device_alloc, freed by the driver) */
    int (*run)(dx_ctx *ctx, const dx_opts *o, const uint8_t *d_in, size_t n,
               uint8_t **d_out, size_t *out_len);
  } dx_tool;

int dx_cli_main(const dx_tool *tool, int argc, char *argv[]);

#endif
