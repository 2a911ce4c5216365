/*
 * oracle/p64_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, single-threaded restatement of the data-parallel hot path of the PVRG-P64 H.261
 * encoder (maikmerten/p64): block-matching motion estimation, the MTYPE decision, prediction
 * (motion compensation, half-vector chroma, loop filter), the integer Chen forward/inverse DCT,
 * quantise / inverse-quantise, zig-zag, CBP + type-4/7 fallback and reconstruction.
 * Every function cites the reference file:line it restates (paths relative to the reference tree).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against the
 * reference's own objects (oracle/_ref/libp64ref.so, built by oracle/build_ref.sh from the
 * reference sources) and tests/test_stream_parity_cpu.py checks whole .p64 streams produced from
 * this oracle + the host bit-stream writer against oracle/_ref/p64_ref{,_fs}.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_ME_TSS  0   /* StepBME, the stock search (me.c:352)                         */
#define ORC_ME_FULL 1   /* FastBME, the commented-out exhaustive search (me.c:351)      */

#define IT_NTSC 0
#define IT_CIF  1
#define IT_QCIF 2

/* MType property tables, p64.c:217-222 */
static const int kQuantM[10]  = {0,1,0,1,0,0,1,0,0,1};
static const int kCbpM[10]    = {0,0,1,1,0,1,1,0,1,1};
static const int kIntraM[10]  = {1,1,0,0,0,0,0,0,0,0};
static const int kMfM[10]     = {0,0,0,0,1,1,1,1,1,1};
static const int kFilterM[10] = {0,0,0,0,0,0,0,1,1,1};
static const int kTcoefM[10]  = {1,1,1,1,0,1,1,0,1,1};

/* zig-zag scan: position of raster coefficient i in the scanned block, transform.c:67-75 */
static const int kZig[64] = {
   0,  1,  5,  6, 14, 15, 27, 28,   2,  4,  7, 13, 16, 26, 29, 42,
   3,  8, 12, 17, 25, 30, 41, 43,   9, 11, 18, 24, 31, 40, 44, 53,
  10, 19, 23, 32, 39, 45, 52, 54,  20, 22, 33, 38, 46, 51, 55, 60,
  21, 34, 37, 47, 50, 56, 59, 61,  35, 36, 48, 49, 57, 58, 62, 63};

/* ------------------------------------------------------------------------------------------ */
/* Motion estimation                                                                           */
/* ------------------------------------------------------------------------------------------ */

/* me.c:92-176 ComputeError without the early exit (a truncated sum is >= MV, so it can never
 * pass the strict `error < MV` tests at me.c:220,300; full sums give identical decisions). */
int orc_sad16(const uint8_t *r, int rs, const uint8_t *c, int cs)
{
  int s = 0;
  for (int y = 0; y < 16; y++)
    for (int x = 0; x < 16; x++) {
      int d = (int)r[y * rs + x] - (int)c[y * cs + x];
      s += d < 0 ? -d : d;
    }
  return s;
}

/* me.c:212-213 / me.c:292-293: strict `<` on the far edge. */
static int legal_pos(int px, int py, int W, int H)
{
  return px >= 0 && px < W - 16 && py >= 0 && py < H - 16;
}

/* One macroblock of MotionEstimation() (me.c:340-363); out = {MX,MY,MV,OMV,VAR,VAROR,MWOR}. */
void orc_me_mb(const uint8_t *ref, const uint8_t *cur, int W, int H, int x0, int y0,
               int mode, int search_limit, int32_t out[7])
{
  const uint8_t *c = cur + y0 * W + x0;
  int mx = 0, my = 0;
  int mv = orc_sad16(ref + y0 * W + x0, W, c, W);     /* (0,0) first: me.c:198-203, 262-271 */
  int omv = mv;
  if (mode == ORC_ME_FULL) {                           /* me.c:206-227 */
    int lo = (-search_limit) / 2, hi = search_limit / 2;
    for (int dx = lo; dx < hi; dx++)
      for (int dy = lo; dy < hi; dy++) {
        int px = x0 + dx, py = y0 + dy;
        if (!legal_pos(px, py, W, H)) continue;
        int e = orc_sad16(ref + py * W + px, W, c, W);
        if (e < mv) { mv = e; mx = dx; my = dy; }
      }
  } else {                                             /* me.c:273-311 */
    int bx = x0, by = y0;
    for (int step = 8; step >= 1; step /= 2) {
      for (int diry = -1; diry <= 1; diry++) {
        int py = by + diry * step, dy = py - y0;
        for (int dirx = -1; dirx <= 1; dirx++) {
          if (!dirx && !diry) continue;
          int px = bx + dirx * step, dx = px - x0;
          if (!legal_pos(px, py, W, H) || dx < -15 || dx > 15 || dy < -15 || dy > 15) continue;
          int e = orc_sad16(ref + py * W + px, W, c, W);
          if (e < mv) { mv = e; mx = dx; my = dy; }
        }
      }
      bx = x0 + mx; by = y0 + my;
    }
  }
  /* statistics on the best-match REFERENCE block: me.c:230-245, 314-329 */
  const uint8_t *b = ref + (y0 + my) * W + (x0 + mx);
  int var = 0, varor = 0, mwor = 0;
  for (int y = 0; y < 16; y++)
    for (int x = 0; x < 16; x++) {
      int rv = b[y * W + x], d = rv - (int)c[y * W + x];
      var += d * d; varor += rv * rv; mwor += rv;
    }
  var /= 256;
  varor = varor / 256 - (mwor / 256) * (mwor / 256);
  out[0] = mx; out[1] = my; out[2] = mv; out[3] = omv; out[4] = var; out[5] = varor; out[6] = mwor;
}

/* MotionEstimation() over a frame, raster MB order (me.c:346-362). out = int32[nmb][7]. */
void orc_me_frame(const uint8_t *ref, const uint8_t *cur, int W, int H, int mode, int search_limit,
                  int32_t *out)
{
  int n = 0;
  for (int y = 0; y < H; y += 16)
    for (int x = 0; x < W; x += 16)
      orc_me_mb(ref, cur, W, H, x, y, mode, search_limit, out + 7 * n++);
}

/* The whole 31x31 SAD surface of one MB (dx,dy in [-15,15]); illegal positions = -1.
 * Not a reference function: a checking aid for the device surface kernel. Index [dy+15][dx+15]. */
void orc_sad_surface(const uint8_t *ref, const uint8_t *cur, int W, int H, int x0, int y0, int32_t *out)
{
  for (int dy = -15; dy <= 15; dy++)
    for (int dx = -15; dx <= 15; dx++) {
      int px = x0 + dx, py = y0 + dy, v = -1;
      if (legal_pos(px, py, W, H) || (dx == 0 && dy == 0))
        v = orc_sad16(ref + py * W + px, W, cur + y0 * W + x0, W);
      out[(dy + 15) * 31 + (dx + 15)] = v;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Chen DCT (chendct.c).  The reference multiplies by `long` constants and stores into int;     */
/* every intermediate fits in 32 bits for the input ranges on the path, but    */
/* 64-bit products are used here anyway so the oracle stays exact for ANY int input.            */
/* ------------------------------------------------------------------------------------------ */
#define MS(e) ((int)(((int64_t)(e)) >> 9))          /* MSCALE, chendct.c:46-53 (arithmetic >>) */
#define K1D4  362LL
#define K1D8  473LL
#define K3D8  196LL
#define K1D16 502LL
#define K3D16 426LL
#define K5D16 284LL
#define K7D16 100LL

static void fdct_1d(const int *s, int sstride, int *d, int dstride, int prescale_shift_left)
{
  int a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3;
  const int *p = s, *q = s + 7 * sstride;
  if (prescale_shift_left) {                 /* column pass: LS(.,2)  chendct.c:119-131 */
    a0 = (p[0] + q[0]) << 2; c3 = (p[0] - q[0]) << 2; p += sstride; q -= sstride;
    a1 = (p[0] + q[0]) << 2; c2 = (p[0] - q[0]) << 2; p += sstride; q -= sstride;
    a2 = (p[0] + q[0]) << 2; c1 = (p[0] - q[0]) << 2; p += sstride; q -= sstride;
    a3 = (p[0] + q[0]) << 2; c0 = (p[0] - q[0]) << 2;
  } else {                                   /* row pass: RS(.,1)  chendct.c:165-172 */
    a0 = (p[0] + q[0]) >> 1; c3 = (p[0] - q[0]) >> 1; p += sstride; q -= sstride;
    a1 = (p[0] + q[0]) >> 1; c2 = (p[0] - q[0]) >> 1; p += sstride; q -= sstride;
    a2 = (p[0] + q[0]) >> 1; c1 = (p[0] - q[0]) >> 1; p += sstride; q -= sstride;
    a3 = (p[0] + q[0]) >> 1; c0 = (p[0] - q[0]) >> 1;
  }
  b0 = a0 + a3; b1 = a1 + a2; b2 = a1 - a2; b3 = a0 - a3;
  d[0 * dstride] = MS(K1D4 * (b0 + b1));
  d[4 * dstride] = MS(K1D4 * (b0 - b1));
  d[2 * dstride] = MS(K3D8 * b2 + K1D8 * b3);
  d[6 * dstride] = MS(K3D8 * b3 - K1D8 * b2);
  b0 = MS(K1D4 * (c2 - c1));
  b1 = MS(K1D4 * (c2 + c1));
  a0 = c0 + b0; a1 = c0 - b0; a2 = c3 - b1; a3 = c3 + b1;
  d[1 * dstride] = MS(K7D16 * a0 + K1D16 * a3);
  d[3 * dstride] = MS(K3D16 * a2 - K5D16 * a1);
  d[5 * dstride] = MS(K3D16 * a1 + K5D16 * a2);
  d[7 * dstride] = MS(K7D16 * a3 - K1D16 * a0);
}

/* ChenDct, chendct.c:97-206 */
void orc_fdct(const int *x, int *y)
{
  int t[64];
  for (int i = 0; i < 8; i++) fdct_1d(x + i, 8, t + i, 8, 1);          /* columns */
  for (int i = 0; i < 8; i++) fdct_1d(t + 8 * i, 1, y + 8 * i, 1, 0);  /* rows (in place in ref) */
  for (int i = 0; i < 64; i++) y[i] = (y[i] < 0 ? y[i] - 4 : y[i] + 4) / 8;   /* chendct.c:204-205 */
}

static void idct_1d(const int *s, int sstride, int *d, int dstride, int prescale)
{
  int a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3;
  int sh = prescale ? 2 : 0;                       /* LS(.,2) on the column pass only */
  b0 = s[0 * sstride] << sh; a0 = s[1 * sstride] << sh; b2 = s[2 * sstride] << sh; a1 = s[3 * sstride] << sh;
  b1 = s[4 * sstride] << sh; a2 = s[5 * sstride] << sh; b3 = s[6 * sstride] << sh; a3 = s[7 * sstride] << sh;
  c0 = MS(K7D16 * a0 - K1D16 * a3);
  c1 = MS(K3D16 * a2 - K5D16 * a1);
  c2 = MS(K3D16 * a1 + K5D16 * a2);
  c3 = MS(K1D16 * a0 + K7D16 * a3);
  a0 = MS(K1D4 * (b0 + b1));
  a1 = MS(K1D4 * (b0 - b1));
  a2 = MS(K3D8 * b2 - K1D8 * b3);
  a3 = MS(K1D8 * b2 + K3D8 * b3);
  b0 = a0 + a3; b1 = a1 + a2; b2 = a1 - a2; b3 = a0 - a3;
  a0 = c0 + c1; a1 = c0 - c1; a2 = c3 - c2; a3 = c3 + c2;
  c0 = a0; c1 = MS(K1D4 * (a2 - a1)); c2 = MS(K1D4 * (a2 + a1)); c3 = a3;
  d[0 * dstride] = b0 + c3; d[1 * dstride] = b1 + c2; d[2 * dstride] = b2 + c1; d[3 * dstride] = b3 + c0;
  d[4 * dstride] = b3 - c0; d[5 * dstride] = b2 - c1; d[6 * dstride] = b1 - c2; d[7 * dstride] = b0 - c3;
}

/* ChenIDct, chendct.c:217-375 */
void orc_idct(const int *x, int *y)
{
  int t[64];
  for (int i = 0; i < 8; i++) idct_1d(x + i, 8, t + i, 8, 1);
  for (int i = 0; i < 8; i++) idct_1d(t + 8 * i, 1, y + 8 * i, 1, 0);
  for (int i = 0; i < 64; i++) y[i] = (y[i] < 0 ? y[i] - 8 : y[i] + 8) / 16;  /* chendct.c:373-374 */
}

/* ------------------------------------------------------------------------------------------ */
/* Quantisers and bounds (transform.c)                                                          */
/* ------------------------------------------------------------------------------------------ */

/* the AC rule shared by CCITTQuantize / CCITTFlatQuantize, transform.c:284-306, 335-348 */
static int quant_ac(int v, int q)
{
  if (q & 1) return v / (2 * q);
  return (v > 0 ? v + 1 : v - 1) / (2 * q);
}
/* CCITTFlatQuantize(m,8,Q) + FlatBoundQuantizeMatrix: transform.c:317-350, 502-518 */
void orc_quant_intra(int *m, int q)
{
  m[0] = m[0] > 0 ? (m[0] + 4) / 8 : (m[0] - 4) / 8;
  for (int i = 1; i < 64; i++) m[i] = quant_ac(m[i], q);
  if (m[0] > 254) m[0] = 254; else if (m[0] < 1) m[0] = 1;
  for (int i = 1; i < 64; i++) { if (m[i] < -127) m[i] = -127; else if (m[i] > 127) m[i] = 127; }
}
/* CCITTQuantize(m,Q,Q) + BoundQuantizeMatrix: transform.c:271-309, 526-537 */
void orc_quant_inter(int *m, int q)
{
  for (int i = 0; i < 64; i++) m[i] = quant_ac(m[i], q);
  for (int i = 0; i < 64; i++) { if (m[i] < -127) m[i] = -127; else if (m[i] > 127) m[i] = 127; }
}
static int iquant_ac(int l, int q)     /* transform.c:374-390, 431-449 */
{
  if (l > 0) return (2 * l + 1) * q - ((q & 1) ? 0 : 1);
  if (l < 0) return (2 * l - 1) * q + ((q & 1) ? 0 : 1);
  return 0;
}
/* ICCITTFlatQuantize(m,8,Q): transform.c:359-392 */
void orc_iquant_intra(int *m, int q)
{
  m[0] = m[0] * 8;
  for (int i = 1; i < 64; i++) m[i] = iquant_ac(m[i], q);
}
/* ICCITTQuantize(m,Q,Q): transform.c:400-451 */
void orc_iquant_inter(int *m, int q)
{
  for (int i = 0; i < 64; i++) m[i] = iquant_ac(m[i], q);
}
/* BoundDctMatrix: transform.c:460-474 (DC clamped from above only) */
void orc_bound_dct(int *m)
{
  if (m[0] > 2047) m[0] = 2047;
  for (int i = 1; i < 64; i++) { if (m[i] < -1023) m[i] = -1023; else if (m[i] > 1023) m[i] = 1023; }
}
/* ZigzagMatrix (scatter) / IZigzagMatrix (gather): transform.c:561-568, 546-553 */
void orc_zigzag(const int *in, int *out)  { for (int i = 0; i < 64; i++) out[kZig[i]] = in[i]; }
void orc_izigzag(const int *in, int *out) { for (int i = 0; i < 64; i++) out[i] = in[kZig[i]]; }

/* ------------------------------------------------------------------------------------------ */
/* Prediction                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* LoadFilterMatrix, io.c:323-372: separable 1-2-1 on one 8x8 block, edges passed through. */
void orc_loop_filter(const uint8_t *p, int stride, int *out)
{
  int t[64];
  for (int i = 0; i < 8; i++) {
    const uint8_t *r = p + i * stride;
    t[8 * i] = r[0] << 2;
    for (int j = 1; j < 7; j++) t[8 * i + j] = r[j - 1] + (r[j] << 1) + r[j + 1];
    t[8 * i + 7] = r[7] << 2;
  }
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) {
      int v;
      if (i == 0 || i == 7) v = t[8 * i + j];
      else v = ((t[8 * i + j] << 1) + t[8 * (i - 1) + j] + t[8 * (i + 1) + j]) >> 2;
      out[8 * i + j] = (v & 2) ? (v >> 2) + 1 : (v >> 2);
    }
}

/* Encoder state: two frame stores (CFS = reference, OFS = being written; p64.c:69-72) and the
 * forced-intra counters (p64.c:213, 1522-1528). */
typedef struct {
  int image_type, W, H, ngob, nmdu;
  uint8_t *ref[3], *out[3];
  uint8_t *last_intra;   /* [ngob*33] */
  int32_t *me;           /* [nmb][7], raster MB order; zero on the first frame (me.c:49-59 globals) */
} orc_enc;

orc_enc *orc_create(int image_type)
{
  orc_enc *e = (orc_enc *)calloc(1, sizeof(orc_enc));
  e->image_type = image_type;
  e->nmdu = 33;
  switch (image_type) {                     /* SetCCITT, p64.c:1476-1514 */
  case IT_NTSC: e->W = 352; e->H = 240; e->ngob = 10; break;
  case IT_CIF:  e->W = 352; e->H = 288; e->ngob = 12; break;
  default:      e->W = 176; e->H = 144; e->ngob = 3;  break;
  }
  for (int j = 0; j < 3; j++) {             /* ClearFS: both stores start zero-filled, p64.c:532-535 */
    int n = j ? (e->W / 2) * (e->H / 2) : e->W * e->H;
    e->ref[j] = (uint8_t *)calloc(n, 1);
    e->out[j] = (uint8_t *)calloc(n, 1);
  }
  e->last_intra = (uint8_t *)calloc(e->ngob * 33, 1);
  e->me = (int32_t *)calloc((e->W / 16) * (e->H / 16) * 7, sizeof(int32_t));
  return e;
}
void orc_destroy(orc_enc *e)
{
  if (!e) return;
  for (int j = 0; j < 3; j++) { free(e->ref[j]); free(e->out[j]); }
  free(e->last_intra); free(e->me); free(e);
}
uint8_t *orc_ref_plane(orc_enc *e, int j) { return e->ref[j]; }
uint8_t *orc_out_plane(orc_enc *e, int j) { return e->out[j]; }
int32_t *orc_me_records(orc_enc *e) { return e->me; }
uint8_t *orc_last_intra(orc_enc *e) { return e->last_intra; }

/* MB (g,m) -> MB column/row: MoveTo, io.c:730-741 */
static void mb_pos(const orc_enc *e, int g, int m, int *col, int *row)
{
  if (e->image_type == IT_QCIF) { *col = m % 11; *row = g * 3 + m / 11; }
  else { *col = (g & 1) * 11 + m % 11; *row = (g >> 1) * 3 + m / 11; }
}

/* GlobalMC (io.c:129-133): ME of the current SOURCE luma against the previous RECONSTRUCTED luma. */
void orc_motion_estimation(orc_enc *e, const uint8_t *src_y, int mode, int search_limit)
{
  orc_me_frame(e->ref[0], src_y, e->W, e->H, mode, search_limit, e->me);
}

/* The MTYPE decision, p64.c:734-773 (double arithmetic exactly as written). Returns MType. */
int orc_decide(int first_frame, int oval, int val, int var, int varor, int last_intra, int force_intra)
{
  int mtype;
  double x = (double)oval, y = (double)val;
  x = x / 256; y = y / 256;
  if (!first_frame) {
    if ((var < 64) || (varor > var)) {
      if ((x < 1.0) || ((x < 3.0) && (y > (x * 0.5))) || (y > (x / 1.1))) mtype = 2;
      else if (var < (double)6) mtype = 5;
      else mtype = 8;
    } else mtype = 0;
    if (force_intra) mtype = 0;            /* `-o < test.intra`: "0 sto MTYPE" (p64.c:758-765) */
  } else mtype = 0;
  if (last_intra > 131) mtype = 0;         /* p64.c:772-773 */
  return mtype;
}

/* Fetch the prediction for block c of the MB at (col,row) under MType `mt` and vector (mvx,mvy):
 * SubOverlay / Sub[F]Compensate / HalfSub[F]Compensate, io.c:142-496.  Chroma vector = MV/2 with C
 * truncation (io.c:268-269). */
static void fetch_pred(const orc_enc *e, int c, int col, int row, int mt, int mvx, int mvy, int *pred)
{
  static const int BJ[6] = {0,0,0,0,1,2}, BV[6] = {0,0,1,1,0,0}, BH[6] = {0,1,0,1,0,0};   /* p64.c:77-79 */
  int j = BJ[c], w = j ? e->W / 2 : e->W;
  int bx = (j ? col : col * 2 + BH[c]) * 8, by = (j ? row : row * 2 + BV[c]) * 8;
  int dx = 0, dy = 0;
  if (kMfM[mt]) { dx = j ? mvx / 2 : mvx; dy = j ? mvy / 2 : mvy; }
  const uint8_t *p = e->ref[j] + (by + dy) * w + bx + dx;
  if (kFilterM[mt]) orc_loop_filter(p, w, pred);
  else for (int i = 0; i < 8; i++) for (int k = 0; k < 8; k++) pred[8 * i + k] = p[i * w + k];
}

/* One macroblock through ReadCompressMDU -> (WriteMDU's inverse half) -> DecodeSaveMDU:
 * p64.c:823-913, 935-959, 971-1013.
 *   src[3]      current source planes
 *   mtype_in    MType chosen by the decision (or 4 with mv=0 when the rate buffer overflowed, p64.c:776-783)
 *   rec[5]      out: final MType, CBP, MVDH, MVDV (as transmitted: zero for non-MC types, marker.c:339-342), UseQuant
 *   levels      out: int[6][64] zig-zag levels (inputbuf, p64.c:167)
 */
void orc_encode_mb(orc_enc *e, const uint8_t *const src[3], int g, int m, int mtype_in, int mvx, int mvy,
                   int quant, int32_t rec[5], int32_t *levels)
{
  static const int BJ[6] = {0,0,0,0,1,2}, BV[6] = {0,0,1,1,0,0}, BH[6] = {0,1,0,1,0,0};
  int col, row, mt = mtype_in, cbp;
  int pred[64], blk[64], coef[64];
  mb_pos(e, g, m, &col, &row);
  for (;;) {
    for (int c = 0; c < 6; c++) {
      int *lv = levels + 64 * c;
      if (!kTcoefM[mt]) { memset(lv, 0, 64 * sizeof(int)); continue; }
      int j = BJ[c], w = j ? e->W / 2 : e->W;
      int bx = (j ? col : col * 2 + BH[c]) * 8, by = (j ? row : row * 2 + BV[c]) * 8;
      const uint8_t *s = src[j] + by * w + bx;
      for (int i = 0; i < 8; i++) for (int k = 0; k < 8; k++) blk[8 * i + k] = s[i * w + k];   /* ReadBlock */
      if (!kIntraM[mt]) {
        fetch_pred(e, c, col, row, mt, mvx, mvy, pred);
        for (int i = 0; i < 64; i++) blk[i] -= pred[i];
      }
      orc_fdct(blk, coef);
      orc_bound_dct(coef);
      if (kIntraM[mt]) orc_quant_intra(coef, quant); else orc_quant_inter(coef, quant);
      orc_zigzag(coef, lv);
    }
    if (!kCbpM[mt]) cbp = 0x3f;
    else {                                  /* p64.c:887-908 */
      int pmask = 0; cbp = 0;
      for (int c = 0; c < 6; c++) {
        int acc = 0;
        for (int i = 0; i < 64; i++) acc += abs(levels[64 * c + i]);
        if (acc && !pmask) pmask |= 1 << (5 - c);
        if (acc > 1) cbp |= 1 << (5 - c);
      }
      if (!cbp) {
        if (pmask) cbp = pmask;
        else { mt = kFilterM[mt] ? 7 : 4; continue; }
      }
    }
    break;
  }
  e->last_intra[g * 33 + m] = kIntraM[mt] ? 0 : (uint8_t)(e->last_intra[g * 33 + m] + 1);   /* p64.c:909-910 */

  /* inverse half + save: p64.c:935-959, 971-1013 */
  for (int c = 0; c < 6; c++) {
    int j = BJ[c], w = j ? e->W / 2 : e->W;
    int bx = (j ? col : col * 2 + BH[c]) * 8, by = (j ? row : row * 2 + BV[c]) * 8;
    if ((cbp & (1 << (5 - c))) && kTcoefM[mt]) {
      orc_izigzag(levels + 64 * c, coef);
      if (kIntraM[mt]) orc_iquant_intra(coef, quant); else orc_iquant_inter(coef, quant);
      orc_idct(coef, blk);
    } else memset(blk, 0, sizeof(blk));
    if (!kIntraM[mt]) {
      fetch_pred(e, c, col, row, mt, mvx, mvy, pred);
      for (int i = 0; i < 64; i++) blk[i] += pred[i];
    }
    uint8_t *o = e->out[j] + by * w + bx;
    for (int i = 0; i < 8; i++) for (int k = 0; k < 8; k++) {
      int v = blk[8 * i + k];
      o[i * w + k] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);      /* BoundIDctMatrix */
    }
  }
  rec[0] = mt; rec[1] = cbp;
  rec[2] = kMfM[mt] ? mvx : 0; rec[3] = kMfM[mt] ? mvy : 0; rec[4] = quant;
}

/* Decision + encode for one MB with the ME records held in the state. `overflow` restates
 * p64.c:776-783 (rate buffer overflow: type 4, zero vector). */
void orc_encode_mb_auto(orc_enc *e, const uint8_t *const src[3], int g, int m, int first_frame, int quant,
                        int force_intra, int overflow, int32_t rec[5], int32_t *levels)
{
  int col, row;
  mb_pos(e, g, m, &col, &row);
  const int32_t *r = e->me + 7 * (row * (e->W / 16) + col);
  int mt = orc_decide(first_frame, r[3], r[2], r[4], r[5], e->last_intra[g * 33 + m], force_intra);
  int mvx = r[0], mvy = r[1];
  if (overflow) { mvx = mvy = 0; mt = 4; }
  orc_encode_mb(e, src, g, m, mt, mvx, mvy, quant, rec, levels);
}

/* A whole frame at one quantiser (fixed-Q mode). recs = int32[ngob*33][5], levels = int32[ngob*33][6][64],
 * both in GOB-major transmission order. */
void orc_encode_frame(orc_enc *e, const uint8_t *sy, const uint8_t *su, const uint8_t *sv, int first_frame,
                      int quant, int me_mode, int search_limit, int force_intra, int32_t *recs, int32_t *levels)
{
  const uint8_t *src[3] = {sy, su, sv};
  if (!first_frame) orc_motion_estimation(e, sy, me_mode, search_limit);     /* p64.c:635-636 */
  for (int g = 0; g < e->ngob; g++)
    for (int m = 0; m < 33; m++)
      orc_encode_mb_auto(e, src, g, m, first_frame, quant, force_intra, 0,
                         recs + 5 * (g * 33 + m), levels + 384 * (g * 33 + m));
}

/* SwapFS(CFS,OFS), p64.c:661 */
void orc_swap(orc_enc *e)
{
  for (int j = 0; j < 3; j++) { uint8_t *t = e->ref[j]; e->ref[j] = e->out[j]; e->out[j] = t; }
}
