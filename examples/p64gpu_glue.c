/* Reference-side binding of maikmerten/p64 to libp64b200.so -- the file a maintainer adds to the reference tree
 * (INTEGRATION.md sections 3, 3b, 3c show it piece by piece; generated from them by tools/make_glue_example.py).  It uses
 * the reference's own headers (globals.h) and globals; tests/test_abi_and_host.py syntax-checks it against them when the
 * reference tree is present. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "globals.h"
#include "p64_b200.h"

extern IMAGE *CImage; extern FRAME *CFrame; extern FSTORE *CFS, *OFS;
extern int ImageType, CurrentFrame, StartFrame, CurrentGOB, CurrentMDU, NumberGOB, NumberMDU;
extern int MType, CBP, MVDH, MVDV, GQuant, UseQuant, SearchLimit, Oracle, Rate;
extern int MeX[], MeY[], MeVal[], MeOVal[], MeVAR[], MeVAROR[], MeMWOR[];
extern unsigned char **LastIntra;
extern int IntraMType[];

static p64b_ctx *ctx; static p64b_mb *mbs; static int8_t *levels; static uint8_t *src, *ovf; static uint8_t q1[1];
static p64b_me *me; static p64b_step step; static int img_t;
int p64gpu_ovf;                                /* set by the MAIN LOOP's overflow branch (p64.c:776-783) for the macroblock at hand */

static void die(void) { BEGIN("p64gpu"); WHEREAMI(); printf("p64gpu: %s\n", p64b_last_error()); exit(ERROR_MEMORY); }

void p64gpu_init(void) {                       /* call once from p64EncodeSequence after ClearFS (p64.c:535) */
  int t = img_t = ImageType == IT_CIF ? P64B_IT_CIF : ImageType == IT_QCIF ? P64B_IT_QCIF : P64B_IT_NTSC;
  if (p64b_ctx_create(&ctx, 0, t, 1)) die();
  src    = p64b_host_alloc(p64b_frame_bytes(t));
  mbs    = p64b_host_alloc(p64b_num_mb(t) * sizeof(p64b_mb));
  levels = p64b_host_alloc(p64b_num_mb(t) * P64B_LEVELS_PER_MB);
  ovf    = calloc(p64b_num_mb(t), 1);
  me     = calloc(p64b_num_mb(t), sizeof(p64b_me));
}

void p64gpu_frame_begin(void) {                /* replaces `if (CurrentFrame!=StartFrame) GlobalMC();` p64.c:635-636 */
  int i, n; unsigned char *d = src;
  const int wh = p64b_width(img_t) * p64b_height(img_t);
  for (i = 0; i < 3; i++) {                    /* ReadIob() already ran: the planes as ReadBlock walks them (io.c:636-645, 793-803) */
    n = i ? wh / 4 : wh;
    memcpy(d, CFrame->Iob[i]->mem->data, n); d += n;
  }
  step.first_frame = CurrentFrame == StartFrame;
#ifdef P64GPU_FULL
  step.me_mode = P64B_ME_FULL;                 /* the FastBME build (me.c:351) */
#else
  step.me_mode = P64B_ME_TSS;
#endif
  step.search_limit = SearchLimit; step.force_intra = 0; step.gquant = GQuant;
  if (p64b_ctx_frame_begin(ctx, &step, src)) die();
  if (!step.first_frame) {                     /* the reference's result arrays (me.c:49-59): the decision code and CallOracle read them */
    if (p64b_ctx_me_records(ctx, 0, me)) die();
    n = (wh / 256);
    for (i = 0; i < n; i++) {
      MeX[i] = me[i].mx; MeY[i] = me[i].my; MeVal[i] = me[i].val; MeOVal[i] = me[i].oval;
      MeVAR[i] = me[i].var; MeVAROR[i] = me[i].varor; MeMWOR[i] = me[i].mwor;
    }
  }
  memset(ovf, 0, NumberGOB * NumberMDU);
}

void p64gpu_gob(void) {                        /* call in p64EncodeGOB right after GQuant is final (p64.c:702) */
  q1[0] = (uint8_t)GQuant;
  if (p64b_ctx_encode_gob(ctx, &step, CurrentGOB, q1, mbs, levels)) die();
}

/* replaces ReadCompressMDU + the inverse half of WriteMDU + DecodeSaveMDU for one MB: sets the globals
   WriteMBHeader()/Encode*() read and fills inputbuf[c][k] with the zig-zag levels; keeps LastIntra (p64.c:909-910) and the
   installed Iob (what Bpos() of the next macroblock's decision code relies on, p64.c:1007-1010) as the reference leaves them */
void p64gpu_mb(int overflow, int inputbuf[10][64]) {
  const p64b_mb *r = &mbs[CurrentMDU]; const int8_t *l = levels + CurrentMDU * P64B_LEVELS_PER_MB; int c, k;
  if (overflow) { MType = 4; CBP = 0x3f; MVDH = MVDV = 0; ovf[CurrentGOB * NumberMDU + CurrentMDU] = 1; }
  else { MType = r->mtype; CBP = r->cbp; MVDH = r->mvx; MVDV = r->mvy; }
  UseQuant = GQuant;
  for (c = 0; c < 6; c++) for (k = 0; k < 64; k++)
    inputbuf[c][k] = overflow ? 0 : (k == 0 && MType < 2) ? (uint8_t)l[64 * c] : l[64 * c + k];   /* intra DC is unsigned */
  if (IntraMType[MType]) LastIntra[CurrentGOB][CurrentMDU] = 0; else LastIntra[CurrentGOB][CurrentMDU]++;
  InstallFS(2, OFS);
}

void p64gpu_frame_end(void) { if (p64b_ctx_frame_end(ctx, Rate ? ovf : 0)) die(); }   /* with SwapFS p64.c:661 */

/* ---- 3b: the device writes the bits too (fixed quantiser) ---- */
static unsigned int pending, pending_len;
extern int FirstFrameBits, NumberOvfl, FrameRate, FrameRateDiv, FrameSkip, QDFact, QOffs;
void p64gpu_frame_bits(void) {                 /* replaces p64EncodeFrame's body between ReadIob() and SwapFS() */
  int64_t t; p64b_bits_out o; size_t i, n; const uint8_t *d;
  /* ... fill src, step as in p64gpu_frame_begin ... */
  if (p64b_ctx_submit_bits(ctx, &step, CurrentFrame % 32, src, &t) || p64b_ctx_wait_bits(ctx, t, &o)) die();
  n = o.nbytes[0]; d = o.data + o.offset[0];
  for (i = 0; i < n; i++) mputv(8, d[i]);      /* whole bytes only (the < 8 pending bits stay on the device), through the
                                                  reference's own writer, so mwtell() and swclose() keep working */
  pending = o.carry[0]; pending_len = o.carry_len[0];
  if (CurrentFrame == StartFrame) FirstFrameBits = (int)o.bit_position[0];
}
/* at end of sequence (p64.c:600-605): mputv(pending >> (32 - pending_len), pending_len); WritePictureHeader(); swclose(); */

/* ---- 3c: rate control without the round trips ---- */
void p64gpu_init_rate(void) {                  /* after p64gpu_init, before the first frame; Rate etc. as set at p64.c:572-590 */
  p64b_rate_control rc = {Rate, FrameRate, FrameRateDiv, FrameSkip, QDFact, QOffs};
  if (p64b_ctx_set_rate_control(ctx, &rc)) die();
}
/* p64gpu_frame_bits() of 3b is unchanged; step.gquant of the FIRST frame carries InitialQuant.  After the call
   GQuant = o.gquant[0]; NumberOvfl = o.overflows[0];  -- ExecuteQuantization(), the overflow branch of the MAIN LOOP and
   the BufferOffset arithmetic at p64.c:670-680 are then dead code on the encode side. */
