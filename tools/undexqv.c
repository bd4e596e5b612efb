/* undexqv -- .dexqv -> .quiva.
 * Same command line, flags and file format as the reference's undexqv (undexqv.c:18-70); the
 * work is done by libdexb200.so on the GPU (see dxcli.h). */
#include "dxcli.h"

static int run(dx_ctx *ctx, const dx_opts *o, const uint8_t *d_in, size_t n,
               uint8_t **d_out, size_t *out_len)
{ size_t cap = 0;
  int rc = dx_undexqv_size_dev(ctx,d_in,n,&cap);
  if (rc != DX_OK) return rc;
  *d_out = (uint8_t *) dx_device_alloc(ctx,cap + 64);
  if (*d_out == NULL) return DX_E_NOMEM;
  return dx_undexqv_dev(ctx,d_in,n,o->upper,*d_out,cap + 64,out_len,NULL,0,0);
}

int main(int argc, char *argv[])
{ static const dx_tool tool =
    { "undexqv", "[-vkU] <path:dexqv> ...", "vkU", 0, ".dexqv", ".quiva",
      { "      -k: do *not* remove the .dexqv file on completion.",
        "      -U: use uppercase letters (default is lower case).", NULL, NULL, NULL }, run };
  return dx_cli_main(&tool,argc,argv);
}
