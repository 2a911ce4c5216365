/* TEST INFRASTRUCTURE ONLY.  The reference's own StatisticsMem() (stat.c:73-130) made callable for tests: the function is
 * `static`, so this translation unit includes the reference source where it lies (-I/root/reference at build time; nothing of
 * it is copied into the repository) and adds a wrapper with external linkage.  Linked into oracle/_ref/libp64ref.so. */
#include "stat.c"

FSTORE *CFS = 0;
STAT *CStat = 0;

/* src, rec: two planes of width x height samples; out[6] = mean, mse, mrsnr, snr, psnr, entropy as the reference computes them */
void ref_statistics_mem(unsigned char *src, unsigned char *rec, int width, int height, double *out)
{
  MEM a, b;
  STAT s;
  a.len = b.len = width * height;
  a.width = b.width = width;
  a.height = b.height = height;
  a.data = src;
  b.data = rec;
  StatisticsMem(&a, &b, &s);
  out[0] = s.mean; out[1] = s.mse; out[2] = s.mrsnr; out[3] = s.snr; out[4] = s.psnr; out[5] = s.entropy;
}
