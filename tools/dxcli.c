/* dxcli.c -- see dxcli.h */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>
#include "dxcli.h"

static const char *Prog;

static void die(dx_ctx *ctx, int rc)
{ if (rc == DX_E_TRUNC)                       /* SYSTEM_READ_ERROR, DB.h:136-139 */
    { fprintf(stderr,"%s: System error, read failed!\n",Prog);
      exit (2);
    }
  fprintf(stderr,"%s: %s\n",Prog,ctx ? dx_strerror(ctx) : "cannot open a CUDA device (no CPU fallback)");
  exit (1);
}

/* whole stream into pinned host memory */
static uint8_t *slurp(dx_ctx *ctx, FILE *f, size_t *n)
{ size_t cap = (size_t) 1 << 26, len = 0;
  uint8_t *buf = (uint8_t *) dx_pinned_alloc(ctx,cap);
  if (buf == NULL) die(ctx,DX_E_NOMEM);
  if (f != stdin && fseeko(f,0,SEEK_END) == 0)
    { off_t sz = ftello(f);
      rewind(f);
      if (sz > 0 && (size_t) sz + 64 > cap)
        { dx_pinned_free(ctx,buf);
          cap = (size_t) sz + 64;
          buf = (uint8_t *) dx_pinned_alloc(ctx,cap);
          if (buf == NULL) die(ctx,DX_E_NOMEM);
        }
    }
  while (1)
    { size_t got = fread(buf+len,1,cap-len,f);
      len += got;
      if (got == 0) break;
      if (len == cap)
        { uint8_t *nb = (uint8_t *) dx_pinned_alloc(ctx,2*cap);
          if (nb == NULL) die(ctx,DX_E_NOMEM);
          memcpy(nb,buf,len);
          dx_pinned_free(ctx,buf);
          buf = nb; cap *= 2;
        }
    }
  *n = len;
  return buf;
}

/* <dir>/<root><ext> from a command-line name, the way PathTo/Root/Catenate do (DB.c:112-181):
   the extension is stripped only if present (compared without case) */
static char *file_name(const char *arg, const char *strip, const char *ext)
{ size_t la = strlen(arg), ls = strlen(strip);
  char *out = (char *) malloc(la + strlen(ext) + 16);
  const char *base = strrchr(arg,'/');
  size_t pre = 0;
  if (base == NULL)                      /* PathTo gives "." for a bare name: "./<root><ext>" */
    { strcpy(out,"./"); pre = 2; base = arg; }
  else
    base = base+1;
  strcpy(out+pre,arg);
  if (strlen(base) > ls && strcasecmp(arg+la-ls,strip) == 0)
    out[pre+la-ls] = '\0';
  strcat(out,ext);
  return out;
}

static char *root_name(const char *arg, const char *strip)
{ char *f = file_name(arg,strip,"");
  char *b = strrchr(f,'/');
  char *r = strdup(b == NULL ? f : b+1);
  free(f);
  return r;
}

int dx_cli_main(const dx_tool *tool, int argc, char *argv[])
{ dx_opts o = { 0, 0, 0, 0, 0, 80 };
  int i, j, k;
  dx_ctx *ctx = NULL;

  Prog = tool->name;
  j = 1;
  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-')
      { if (tool->has_width && argv[i][1] == 'w')                    /* ARG_NON_NEGATIVE */
          { char *eptr;
            o.width = (int) strtol(argv[i]+2,&eptr,10);
            if (*eptr != '\0' || argv[i][2] == '\0')
              { fprintf(stderr,"%s: -%c '%s' argument is not an integer\n",Prog,argv[i][1],argv[i]+2);
                exit (1);
              }
            if (o.width < 0)
              { fprintf(stderr,"%s: %s must be non-negative (%d)\n",Prog,"Line width",o.width);
                exit (1);
              }
            if (o.width == 0)      /* the reference loops forever on -w0 (undexta.c:265) */
              { fprintf(stderr,"%s: Line width must be positive (0)\n",Prog);
                exit (1);
              }
            continue;
          }
        for (k = 1; argv[i][k] != '\0'; k++)                          /* ARG_FLAGS */
          { if (strchr(tool->flags,argv[i][k]) == NULL)
              { fprintf(stderr,"%s: -%c is an illegal option\n",Prog,argv[i][k]);
                exit (1);
              }
            switch (argv[i][k])
              { case 'v': o.verbose = 1; break;
                case 'k': o.keep = 1; break;
                case 'i': o.pipe = 1; break;
                case 'U': o.upper = 1; break;
                case 'l': o.lossy = 1; break;
              }
          }
      }
    else
      argv[j++] = argv[i];
  argc = j;

  if ((o.pipe && argc > 1) || (!o.pipe && argc <= 1))
    { fprintf(stderr,"Usage: %s %s\n",Prog,tool->usage);
      fprintf(stderr,"\n");
      for (k = 0; k < 5 && tool->help[k] != NULL; k++)
        fprintf(stderr,"%s\n",tool->help[k]);
      exit (1);
    }
  if (o.pipe)
    { o.keep = 1; argc = 2; }

  for (i = 1; i < argc; i++)
    { char *src = NULL, *dst = NULL, *root;
      FILE *in, *out = NULL;
      uint8_t *h_in, *h_out, *d_in, *d_out = NULL;
      size_t n, m = 0;
      int rc;

      if (o.pipe)
        { in = stdin; out = stdout;
          root = strdup("Standard Input");
        }
      else
        { src  = file_name(argv[i],tool->src_ext,tool->src_ext);
          dst  = file_name(argv[i],tool->src_ext,tool->dst_ext);
          root = root_name(argv[i],tool->src_ext);
          if ((in = fopen(src,"r")) == NULL)
            { fprintf(stderr,"%s: Cannot open %s for 'r'\n",Prog,src); exit (1); }
        }
      /* the device is opened once a source is: a missing file is reported the way the reference
         reports it, and no output file is created when there is no GPU to fill it */
      if (ctx == NULL)
        { int dev = 0;
          const char *e = getenv("DEXB200_DEVICE");
          if (e != NULL) dev = atoi(e);
          if (dx_open(dev,&ctx) != DX_OK) die(NULL,DX_E_NOGPU);
        }
      if (!o.pipe && (out = fopen(dst,"w")) == NULL)
        { fprintf(stderr,"%s: Cannot open %s for 'w'\n",Prog,dst); exit (1); }
      if (o.verbose)
        { fprintf(stderr,"Processing '%s' ...\n",root); fflush(stderr); }

      h_in = slurp(ctx,in,&n);
      d_in = (uint8_t *) dx_device_alloc(ctx,n + 64);
      if (d_in == NULL) die(ctx,DX_E_NOMEM);
      if ((rc = dx_h2d(ctx,d_in,h_in,n)) != DX_OK) die(ctx,rc);
      if ((rc = tool->run(ctx,&o,d_in,n,&d_out,&m)) != DX_OK) die(ctx,rc);
      h_out = (uint8_t *) dx_pinned_alloc(ctx,m + 1);
      if (h_out == NULL) die(ctx,DX_E_NOMEM);
      if ((rc = dx_d2h(ctx,h_out,d_out,m)) != DX_OK) die(ctx,rc);
      if ((rc = dx_sync(ctx)) != DX_OK) die(ctx,rc);
      if (m > 0 && fwrite(h_out,1,m,out) != m)
        { fprintf(stderr,"%s: System error, write failed!\n",Prog); exit (2); }

      /* the source is removed only after the output is completely written */
      if (!o.pipe)
        { fclose(in);
          if (fclose(out) != 0)
            { fprintf(stderr,"%s: System error, write failed!\n",Prog); exit (2); }
          if (!o.keep) unlink(src);
        }
      else
        fflush(out);
      dx_pinned_free(ctx,h_in); dx_pinned_free(ctx,h_out);
      dx_device_free(ctx,d_in); dx_device_free(ctx,d_out);
      free(src); free(dst); free(root);
      if (o.verbose)
        { fprintf(stderr,"Done\n"); fflush(stderr); }
    }
  dx_close(ctx);
  return 0;
}
