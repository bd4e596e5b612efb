#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "dexb200.h"
int main(void)
{ dx_qv_stats *st = malloc(sizeof(*st)); dx_qv_coding *cd = malloc(sizeof(*cd));
  uint8_t *o = malloc(200000); char prefix[100001]; size_t w, used; long ok = 0, bad = 0, rt = 0;
  srand(11);
  for (int t = 0; t < 3000; t++)
    { memset(st,0,sizeof(*st));
      int mode = t % 5;
      for (int k = 0; k < 6; k++)
        { int nsym = 1 + rand() % 256;
          for (int j = 0; j < nsym; j++)
            { int x = rand() % 256;
              uint64_t v = (mode == 0) ? 1 + rand() % 4 : (mode == 1) ? (uint64_t) rand() * rand()
                         : (mode == 2) ? (uint64_t) pow(2.0,(rand() % 4000)/100.0) + 1 : (mode == 3) ? 1 : (uint64_t) 1 << (rand() % 62);
              st->hist[k][x] = v;
            }
        }
      st->totchar = (uint64_t) rand() * 1000; st->nentries = 10;
      st->delchar = (rand() & 1) ? rand() % 128 : -1; st->subchar = (rand() & 1) ? rand() % 128 : -1;
      if (dx_qv_make_coding(st,rand() & 1,cd) != 0) { bad++; continue; }
      ok++;
      if (dx_qv_write_coding(cd,"m140913",7,o,200000,&w) != 0) continue;
      uint8_t *c = malloc(w); memcpy(c,o,w);
      dx_qv_coding *c2 = malloc(sizeof(*c2));
      if (dx_qv_read_coding(c,w,c2,prefix,sizeof(prefix),&used) == 0 && used == w) rt++;
      free(c2); free(c);
    }
  printf("coded %ld refused %ld header round trips %ld\n",ok,bad,rt);
  return 0;
}
