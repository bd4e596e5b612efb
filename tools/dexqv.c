/* dexqv -- .quiva -> .dexqv (per-file Huffman + run-length coding of the five QV streams).
 * Same command line, flags and file format as the reference's dexqv (dexqv.c:22-53); the work is
 * done by libdexb200.so on the GPU (see dxcli.h). */
#include "dxcli.h"

static int run(dx_ctx *ctx, const dx_opts *o, const uint8_t *d_in, size_t n,
               uint8_t **d_out, size_t *out_len)
{ size_t cap = 3*n + 200000;            /* worst case: every symbol escaped (24 bits) */
  *d_out = (uint8_t *) dx_device_alloc(ctx,cap);
  if (*d_out == NULL) return DX_E_NOMEM;
  return dx_dexqv_dev(ctx,d_in,n,o->lossy,*d_out,cap,out_len);
}

int main(int argc, char *argv[])
{ static const dx_tool tool =
    { "dexqv", "[-vkl] <path:quiva> ...", "vkl", 0, ".quiva", ".dexqv",
      { "      -k: do *not* remove the .quiva file on completion.",
        "      -l: use lossy compression (not recommended).", NULL, NULL, NULL }, run };
  return dx_cli_main(&tool,argc,argv);
}
