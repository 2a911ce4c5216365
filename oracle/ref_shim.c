/* TEST INFRASTRUCTURE ONLY. Globals that the reference's io.c declares `extern` (io.c:62-70) and its
 * p64.c would define; linked into oracle/_ref/libp64ref.so so the reference's own hot-path objects
 * (me.c, mem.c, chendct.c, transform.c, io.c) load stand-alone for function-level checks. */
void *CFrame = 0;
int ImageType = 0;
int MVDH = 0;
int MVDV = 0;

/* What the reference encoder sees of a Y4M file (ReadIob, io.c:636-645, through the reference's OWN reader
 * vidinput.c / y4m_input.c, chroma conversion included): per frame the luma plane, then the first (w/2)*(h/2) bytes of
 * each chroma plane -- ReadBlock (io.c:793-803) walks them with the CIF/QCIF stride, whatever the plane's real shape.
 * Returns the number of frames copied into out[max_frames][w*h*3/2], or -1. */
#include <stdio.h>
#include <string.h>
#include "vidinput.h"
int ref_y4m_frames(const char *path, int max_frames, int w, int h, unsigned char *out)
{
  video_input vid;
  video_input_ycbcr frame;
  char tag[5];
  int n = 0, csz = (w / 2) * (h / 2);
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  if (video_input_open(&vid, f) < 0) { fclose(f); return -1; }
  while (n < max_frames && video_input_fetch_frame(&vid, frame, tag) > 0) {
    unsigned char *o = out + (size_t)n * (w * h + 2 * csz);
    memcpy(o, frame[0].data, (size_t)w * h);
    memcpy(o + w * h, frame[1].data, csz);
    memcpy(o + w * h + csz, frame[2].data, csz);
    n++;
  }
  video_input_close(&vid);
  return n;
}
