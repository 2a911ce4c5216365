/* Drop-in build of the reference: maikmerten/p64's OWN main(), flag parsing, Y4M reader and stream writer, with the body of
 * p64EncodeFrame() (p64.c:633-652: GlobalMC, WritePictureHeader, the GOB loop = decision, ReadCompressMDU, WriteMDU,
 * DecodeSaveMDU, and under -r ExecuteQuantization + the overflow override) replaced by ONE call into libp64b200.so per frame.
 * oracle/build_ref.sh compiles the reference sources where they lie, with p64.c passed through a handful of sed edits
 * (into a temp file outside the repository) that insert the four calls below, and links this file + the library into
 * oracle/_ref/p64_gpu (stock three-step search) and p64_gpu_fs (-DP64GPU_FULL: FastBME, me.c:351).
 * tests/test_dropin.py runs those binaries on a B200 and compares their .p64 output with the unmodified reference's. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "globals.h"
#include "p64_b200.h"

extern FRAME *CFrame;
extern int ImageType, CurrentFrame, StartFrame, GQuant, SearchLimit, Rate, InitialQuant;
extern int FrameRate, FrameRateDiv, FrameSkip, QDFact, QOffs, NumberOvfl;

static p64b_ctx *ctx;
static uint8_t *src;
static unsigned int pending, pending_len;
static int image_type, started;

static void die(void) { BEGIN("p64gpu"); WHEREAMI(); printf("p64gpu: %s\n", p64b_last_error()); exit(ERROR_MEMORY); }

void p64gpu_init(void) {                        /* after swopen() in p64EncodeSequence: MakeFstore/InitFS/ClearFS live in HBM */
  image_type = ImageType == IT_CIF ? P64B_IT_CIF : ImageType == IT_QCIF ? P64B_IT_QCIF : P64B_IT_NTSC;
  if (p64b_ctx_create(&ctx, 0, image_type, 1)) die();
  src = (uint8_t *)p64b_host_alloc((size_t)p64b_frame_bytes(image_type));
  if (!src) die();
}

void p64gpu_init_rate(void) {                   /* after GQuant=MQuant=InitialQuant (p64.c:590): Rate, QDFact, QOffs are final */
  p64b_rate_control rc;
  if (!Rate) return;
  memset(&rc, 0, sizeof rc);
  rc.rate = Rate; rc.frame_rate = FrameRate; rc.frame_rate_div = FrameRateDiv; rc.frame_skip = FrameSkip;
  rc.qdfact = QDFact; rc.qoffs = QOffs;
  if (p64b_ctx_set_rate_control(ctx, &rc)) die();
}

void p64gpu_frame_bits(void) {                  /* replaces GlobalMC + WritePictureHeader + the p64EncodeGOB loop */
  p64b_step step; p64b_bits_out o; int64_t t; size_t i, n; const uint8_t *d; uint8_t *dst = src;
  const size_t wh = (size_t)p64b_width(image_type) * p64b_height(image_type);
  /* ReadIob() already ran: the planes as ReadBlock would walk them (io.c:636-645, 793-803) */
  memcpy(dst, CFrame->Iob[0]->mem->data, wh); dst += wh;
  memcpy(dst, CFrame->Iob[1]->mem->data, wh / 4); dst += wh / 4;
  memcpy(dst, CFrame->Iob[2]->mem->data, wh / 4);
  memset(&step, 0, sizeof step);
  step.first_frame = CurrentFrame == StartFrame;
#ifdef P64GPU_FULL
  step.me_mode = P64B_ME_FULL;
#else
  step.me_mode = P64B_ME_TSS;
#endif
  step.search_limit = SearchLimit; step.gquant = GQuant;
  if (p64b_ctx_submit_bits(ctx, &step, CurrentFrame % 32, src, &t) || p64b_ctx_wait_bits(ctx, t, &o)) die();
  n = o.nbytes[0]; d = o.data + o.offset[0];
  for (i = 0; i < n; i++) mputv(8, d[i]);       /* whole bytes, through the reference's own writer (stream.c:193) */
  pending = o.carry[0]; pending_len = o.carry_len[0];
  GQuant = (int)o.gquant[0]; NumberOvfl = (int)o.overflows[0];
  started = 1;
}

void p64gpu_flush(void) {                       /* before the trailing WritePictureHeader (p64.c:601-604) */
  if (started && pending_len) mputv((int)pending_len, (int)(pending >> (32 - pending_len)));
  p64b_ctx_destroy(ctx); ctx = 0;
}
