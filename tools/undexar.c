/* undexar -- .dexar -> .arrow.
 * Same command line, flags and file format as the reference's undexar (undexar.c:19-91); the
 * work is done by libdexb200.so on the GPU (see dxcli.h). */
#include "dxcli.h"

static int run(dx_ctx *ctx, const dx_opts *o, const uint8_t *d_in, size_t n,
               uint8_t **d_out, size_t *out_len)
{ size_t cap = 0;
  int rc = dx_undexta_size_dev(ctx,DX_ARROW,d_in,n,o->width,&cap);
  if (rc != DX_OK) return rc;
  *d_out = (uint8_t *) dx_device_alloc(ctx,cap + 64);
  if (*d_out == NULL) return DX_E_NOMEM;
  return dx_undexta_dev(ctx,DX_ARROW,d_in,n,o->width,0,*d_out,cap + 64,out_len);
}

int main(int argc, char *argv[])
{ static const dx_tool tool =
    { "undexar", "[-vk] [-w<int(80)>] ( -i | <path:dexar> ... )", "vki", 1, ".dexar", ".arrow",
      { "      -i: source is on standard input.",
        "      -k: do *not* remove the .dexar file on completion.",
        "      -w: line width for arrow lines.", NULL, NULL }, run };
  return dx_cli_main(&tool,argc,argv);
}
