// Host check of dextractor_b200/csrc/dx_runmask.cuh: the bit-mask run-length counts of a line against
// a direct restatement of Histogram_Runs (reference QV.c:709-724), over random lines, run densities,
// lengths around the span and bucket boundaries, and every alignment of the line against the spans.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "dx_runmask.cuh"

static void direct(uint64_t run[256], const uint8_t *s, int rlen, int rc)
{ int k = 0;
  while (k < rlen)
    { const int h = k;
      while (k < rlen && s[k] == rc) k++;
      run[k-h >= 255 ? 255 : k-h] += 1;
      if (k < rlen) k++;
    }
}

int main(void)
{ static uint8_t line[70000];
  const int lens[] = { 1, 2, 3, 31, 32, 33, 63, 64, 65, 95, 96, 97, 254, 255, 256, 257, 287, 288, 289, 511, 512, 513, 1000, 4097, 65537 };
  const double dens[] = { 0.0, 0.02, 0.5, 0.88, 0.97, 0.995, 0.9995, 1.0 };
  long checked = 0;
  srand(3);
  for (unsigned li = 0; li < sizeof(lens)/sizeof(lens[0]); li++)
    for (unsigned di = 0; di < sizeof(dens)/sizeof(dens[0]); di++)
      for (int rep = 0; rep < 3; rep++)
        { const int rlen = lens[li];
          for (int i = 0; i < rlen; i++)
            line[i] = (rand() / (double) RAND_MAX < dens[di]) ? 50 : (uint8_t) (33 + rand() % 17);
          if (rep == 1 && rlen > 1) line[rlen-1] = 40;           // ends in an item
          if (rep == 2) line[0] = 40;                            // starts with an item
          uint64_t want[256]; memset(want,0,sizeof(want));
          direct(want,line,rlen,50);
          for (int skew = 0; skew < 32; skew++)
            { uint64_t got[256]; memset(got,0,sizeof(got));
              dx_runs_line_model(got,line,rlen,50,skew);
              if (memcmp(got,want,sizeof(want)) != 0)
                { printf("MISMATCH rlen %d density %g rep %d skew %d\n",rlen,dens[di],rep,skew);
                  for (int g = 0; g < 256; g++)
                    if (got[g] != want[g]) printf("  run[%d] got %llu want %llu\n",g,(unsigned long long) got[g],(unsigned long long) want[g]);
                  return 1;
                }
              checked++;
            }
        }
  printf("ok %ld line/alignment cases\n",checked);
  return 0;
}
