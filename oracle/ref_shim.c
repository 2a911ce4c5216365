/* TEST INFRASTRUCTURE ONLY. Globals that the reference's io.c declares `extern` (io.c:62-70) and its
 * p64.c would define; linked into oracle/_ref/libp64ref.so so the reference's own hot-path objects
 * (me.c, mem.c, chendct.c, transform.c, io.c) load stand-alone for function-level checks. */
void *CFrame = 0;
int ImageType = 0;
int MVDH = 0;
int MVDV = 0;
