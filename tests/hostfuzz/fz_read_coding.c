#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "dexb200.h"
/* usage: asan_fz header.bin : truncations + random corruptions through dx_qv_read_coding, each on an
   exactly-sized heap copy so that ASan sees any read past the end */
int main(int argc, char **argv)
{ FILE *f = fopen(argv[1],"rb"); uint8_t *h = malloc(1 << 20); size_t n = fread(h,1,1 << 20,f); fclose(f);
  dx_qv_coding *cd = malloc(sizeof(dx_qv_coding)); char prefix[100001]; size_t used;
  long ok = 0, bad = 0;
  for (size_t k = 0; k <= n; k++)
    { uint8_t *c = malloc(k ? k : 1); memcpy(c,h,k);
      if (dx_qv_read_coding(c,k,cd,prefix,sizeof(prefix),&used) == 0) ok++; else bad++;
      free(c);
    }
  srand(7);
  for (int t = 0; t < 20000; t++)
    { uint8_t *c = malloc(n); memcpy(c,h,n);
      int m = 1 + rand() % 3;
      for (int j = 0; j < m; j++) c[rand() % n] = (uint8_t) rand();
      if (dx_qv_read_coding(c,n,cd,prefix,sizeof(prefix),&used) == 0)
        { ok++;
          /* a header that parses must also serialise without leaving its buffers */
          uint8_t *o = malloc(200000); size_t w = 0;
          dx_qv_write_coding(cd,prefix,(int) strlen(prefix),o,200000,&w); free(o);
        }
      else bad++;
      free(c);
    }
  printf("ok %ld refused %ld\n",ok,bad);
  return 0;
}
