/*
 * p64_b200.h -- C ABI of the B200-native hot path for the PVRG-P64 H.261 encoder (maikmerten/p64).
 *
 * Plain C, pointers and sizes only.  Three groups of entry points:
 *
 *  (1) p64b_ctx_*   the device context: frame stores resident in HBM for N independent streams and the
 *                   sm_100a kernels (motion estimation, MTYPE decision, prediction + loop filter,
 *                   Chen DCT, quantise, inverse quantise, Chen IDCT, reconstruct).  This is what the
 *                   reference's per-frame work in p64EncodeFrame()/p64EncodeGOB() binds to.
 *  (2) p64b_bits_*  the sequential host side that stays on the CPU: H.261 picture/GOB/MB headers and the
 *                   run-level VLC, consuming the per-macroblock records the device returns.
 *  (3) p64b_enc_*   the sequence driver: (1)+(2)+ the reference's rate control for a batch of streams;
 *                   what the `p64b` command line and the Python mirror call.
 *
 * All references "file:line" are into the reference tree (maikmerten/p64).
 * Functions returning int return 0 on success and a negative P64B_E* code on failure;
 * p64b_last_error() gives the message.  There is NO CPU fallback: without a CUDA device every
 * p64b_ctx_ and p64b_enc_ call that needs one fails with P64B_ECUDA.
 */
#ifndef P64_B200_H
#define P64_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P64B_VERSION 2

/* image types: globals.h IT_NTSC/IT_CIF/IT_QCIF, selected by -NTSC/-CIF/-QCIF (p64.c:281-292) */
#define P64B_IT_NTSC 0 /* 352x240, 10 GOBs */
#define P64B_IT_CIF  1 /* 352x288, 12 GOBs */
#define P64B_IT_QCIF 2 /* 176x144,  3 GOBs */

/* motion search: the stock three-step search (StepBME, me.c:255-330, called at me.c:352) or the
 * exhaustive search (FastBME, me.c:187-246, the commented-out call at me.c:351) */
#define P64B_ME_TSS  0
#define P64B_ME_FULL 1

#define P64B_EINVAL (-1)
#define P64B_ECUDA  (-2)
#define P64B_ENOMEM (-3)
#define P64B_EIO    (-4)

#define P64B_MDU_PER_GOB 33 /* NumberMDU, p64.c:1478 */

/* Y4M chroma types the reference's reader accepts (y4m_input.c:587-655) */
#define P64B_CHROMA_420JPEG  0 /* "420", "420jpeg": no conversion                                   */
#define P64B_CHROMA_420MPEG2 1 /* horizontal re-siting filter (y4m_input.c:195-229)                 */
#define P64B_CHROMA_420PALDV 2 /* horizontal + vertical re-siting (y4m_input.c:274-376)             */
#define P64B_CHROMA_422      3 /* the 420mpeg2 filter on full-height planes (y4m_input.c:617-624)   */
#define P64B_CHROMA_411      4 /* 4:1:1 -> 4:2:2 (y4m_input.c:417-459)                              */
#define P64B_CHROMA_444      5 /* no conversion                                                     */
#define P64B_CHROMA_444ALPHA 6 /* no conversion, alpha plane dropped                                */
#define P64B_CHROMA_MONO     7 /* chroma = 128 (y4m_input.c:463-471)                                */

/* One macroblock as the host VLC needs it (the globals WriteMBHeader()/WriteMDU() read:
 * MType, CBP, MVDH, MVDV, UseQuant -- p64.c:85-98, marker.c:288-354).  8 bytes. */
typedef struct p64b_mb {
  uint8_t mtype;   /* final MType 0..9 after the CBP / type-4 / type-7 fallback (p64.c:887-908)       */
  uint8_t cbp;     /* coded block pattern, bit (5-c) for block c; 0x3f for non-CBP types              */
  int8_t  mvx;     /* MVDH as transmitted: the ME vector for MC types, 0 otherwise (marker.c:339-342) */
  int8_t  mvy;     /* MVDV                                                                            */
  uint8_t quant;   /* UseQuant (p64.c:832-837)                                                        */
  uint8_t nzmask;  /* bit (5-c) set iff block c has any non-zero level (host fast path; not in ref)   */
  uint16_t reserved;
} p64b_mb;

/* Zig-zag levels: int8 [6][64] per macroblock in transmission order (inputbuf[c][k], p64.c:167).
 * AC levels are in [-127,127]; the DC level of an INTRA block is in [1,254] and is stored as uint8
 * in the same byte. */
#define P64B_LEVELS_PER_MB 384

/* Motion-estimation record, the reference's MeX/MeY/MeVal/MeOVal/MeVAR/MeVAROR/MeMWOR (me.c:49-59). */
typedef struct p64b_me {
  int32_t mx, my, val, oval, var, varor, mwor, pad;
} p64b_me;

/* What one frame step needs besides pixels. */
typedef struct p64b_step {
  int32_t first_frame;  /* 1: CurrentFrame==StartFrame: no ME, every MB intra (p64.c:635, 770-771)       */
  int32_t me_mode;      /* P64B_ME_TSS | P64B_ME_FULL                                                   */
  int32_t search_limit; /* -i SearchLimit (p64.c:341-344); FULL searches [-limit/2, limit/2) (me.c:206)  */
  int32_t force_intra;  /* stand-in for `-o < test.intra` ("0 sto MTYPE"): decision forced to type 0     */
  int32_t gquant;       /* GQUANT for every GOB when `quant` is NULL (fixed-Q, -q)                      */
  int32_t reserved[3];
} p64b_step;

const char *p64b_last_error(void);
int p64b_version(void);

/* geometry (SetCCITT, p64.c:1476-1514) */
int p64b_width(int image_type);
int p64b_height(int image_type);
int p64b_frame_bytes(int image_type); /* planar Y,U,V 4:2:0, no padding (mem.h:37-42) */
int p64b_num_gob(int image_type);
int p64b_num_mb(int image_type);      /* NumberGOB*33 */

/* ---------------------------------------------------------------------------------------------
 * (1) device context
 * ------------------------------------------------------------------------------------------- */
typedef struct p64b_ctx p64b_ctx;

/* Replaces MakeFstore/InitFS/ClearFS (p64.c:529-535): two zero-filled frame stores per stream, in HBM. */
int p64b_ctx_create(p64b_ctx **out, int device, int image_type, int n_streams);
void p64b_ctx_destroy(p64b_ctx *ctx);
int p64b_ctx_streams(const p64b_ctx *ctx);
/* Use an existing CUDA stream (cudaStream_t) for every kernel and copy; NULL = the context's own. */
int p64b_ctx_set_cuda_stream(p64b_ctx *ctx, void *cuda_stream);
/* Pinned host memory for source/record/level buffers (cudaHostAlloc); plain malloc'd buffers also work. */
void *p64b_host_alloc(size_t bytes);
void p64b_host_free(void *p);

/* Ingest.  By default `src` frames are 4:2:0 with "jpeg" chroma siting.  After
 * p64b_ctx_set_input_chroma() (before the first frame) every host `src` argument of this context is instead the
 * unconverted Y4M frame payload, [n_streams][p64b_raw_frame_bytes(image_type, chroma)], and the chroma conversion the
 * reference's reader applies on the CPU (y4m_input.c:195-545) runs on the device after the upload.  What the encoder
 * reads of 4:2:2 / 4:1:1 / 4:4:4 input is what the reference reads (the first (W/2)*(H/2) bytes of each converted chroma
 * plane: ReadIob io.c:636-645 + ReadBlock io.c:793-803), so streams stay byte-identical.  Device-pointer entry points
 * (*_dev) always take converted 4:2:0 frames. */
int p64b_raw_frame_bytes(int image_type, int chroma);
int p64b_ctx_set_input_chroma(p64b_ctx *ctx, int chroma);
/* Upload + convert only: raw [n_streams][raw_frame_bytes] -> out420 [n_streams][frame_bytes] (what ReadIob would install). */
int p64b_ctx_convert_frames(p64b_ctx *ctx, const uint8_t *raw, uint8_t *out420);

/* One frame step for every stream, host buffers in and out.  Replaces the body of p64EncodeFrame()
 * between ReadIob() and SwapFS() (p64.c:633-661) in fixed-quantiser mode: GlobalMC/MotionEstimation
 * (io.c:129, me.c:340), the MTYPE decision (p64.c:734-773), ReadCompressMDU (p64.c:823-913), the
 * inverse half of WriteMDU (p64.c:952-957), DecodeSaveMDU (p64.c:971-1013) and SwapFS (p64.c:661).
 *   src     [n_streams][frame_bytes]        current source frames (ReadIob, io.c:613-646)
 *   mbs     [n_streams][num_mb]             GOB-major (transmission) order
 *   levels  [n_streams][num_mb][6][64]
 */
int p64b_ctx_encode_frames(p64b_ctx *ctx, const p64b_step *step, const uint8_t *src, p64b_mb *mbs,
                           int8_t *levels);
/* The same step, pipelined: submit() returns as soon as the work is enqueued (upload on a copy stream, kernels on
 * the context's stream, download on a second copy stream) so that consecutive steps overlap their PCIe traffic with
 * each other's kernels; wait(ticket) blocks until that step's mbs/levels are complete in the caller's buffers.
 * src/mbs/levels must stay valid (and src unmodified) until wait() returns; use pinned memory (p64b_host_alloc).
 * At most 3 steps may be in flight; p64b_ctx_encode_frames() == submit + wait. */
int p64b_ctx_submit(p64b_ctx *ctx, const p64b_step *step, const uint8_t *src, p64b_mb *mbs, int8_t *levels,
                    int64_t *ticket);
int p64b_ctx_wait(p64b_ctx *ctx, int64_t ticket);
/* Same with device pointers (inputs already resident in HBM, outputs left there). */
int p64b_ctx_encode_frames_dev(p64b_ctx *ctx, const p64b_step *step, const uint8_t *src_dev,
                               p64b_mb *mbs_dev, int8_t *levels_dev);

/* Device-side entropy coding: the same fixed-quantiser frame step, but the picture / GOB /
 * macroblock headers and the run-level VLC (marker.c:103-354, codec.c:96-355) also run on the device, and what comes
 * back is every stream's next whole bytes of the H.261 stream -- the host only appends them.  The stream's pending
 * bits (< 8) stay on the device between frames.  submit_bits() enqueues like p64b_ctx_submit() (at most 3 steps in
 * flight); wait_bits() blocks until the step's output is in pinned host memory owned by the context, valid until the
 * third following submit.  temporal_reference = CurrentFrame % 32 (p64.c:637).  Records and levels are not downloaded. */
typedef struct p64b_bits_out {
  const uint8_t  *data;          /* packed chunks of all streams                                              */
  const uint32_t *offset;        /* [n_streams+1] byte offset of stream s's chunk in data; [n_streams] = total */
  const uint32_t *nbytes;        /* [n_streams] whole bytes of stream s's chunk                               */
  const uint32_t *carry;         /* [n_streams] pending bits after this frame, left-aligned in the word       */
  const uint32_t *carry_len;     /* [n_streams] how many (0..7)                                               */
  const uint64_t *bit_position;  /* [n_streams] bits written so far (mwtell, stream.c:233-238)                */
  size_t total_bytes;
  size_t downloaded_bytes;       /* what the device-to-host copies of this step moved (header + budgeted data)  */
  const uint32_t *gquant;        /* [n_streams] GQuant after this frame (p64.c:89; varies under rate control)   */
  const uint32_t *overflows;     /* [n_streams] NumberOvfl so far (p64.c:780; 0 without rate control)          */
} p64b_bits_out;
int p64b_ctx_submit_bits(p64b_ctx *ctx, const p64b_step *step, int temporal_reference, const uint8_t *src,
                         int64_t *ticket);
int p64b_ctx_wait_bits(p64b_ctx *ctx, int64_t ticket, p64b_bits_out *out);
/* The frame step of p64b_ctx_submit_bits with the source frames already on the device and the output left there (fixed
 * quantiser only): motion estimation, macroblock kernel and the headers + VLC kernels, nothing copied.  For measurements
 * that include the entropy coding but not the host link (bench.py `value_with_vlc`); asynchronous on the context's stream.
 * Not to be mixed with steps still in flight from p64b_ctx_submit_bits (it uses pipeline slot 0's buffers). */
int p64b_ctx_encode_bits_dev(p64b_ctx *ctx, const p64b_step *step, int temporal_reference, const uint8_t *src_dev);

/* Rate control on the device (-r).  Once configured (before the first frame), p64b_ctx_submit_bits()
 * runs the reference's buffer model per stream ON THE DEVICE: motion estimation once per frame, then GOB by GOB
 * quantise..reconstruct with the stream's own GQUANT, entropy-code, count the bits, pick the next GOB's GQUANT
 * (BufferContents / ExecuteQuantization, p64.c:233-237, 458-481, 697-702) and apply the per-macroblock overflow override
 * (p64.c:776-783) -- the dependency "quantiser of GOB g <- bits of GOBs < g" is kept, it just no longer crosses PCIe.
 * step->gquant of the FIRST frame is the initial quantiser (InitialQuant, p64.c:574-590); later frames ignore it. */
typedef struct p64b_rate_control {
  int32_t rate;            /* Rate, bits/s (-r); 0 = off                                   */
  int32_t frame_rate;      /* FrameRate / FrameRateDiv (-f, p64.c:128-129)                 */
  int32_t frame_rate_div;
  int32_t frame_skip;      /* FrameSkip (-k)                                               */
  int32_t qdfact;          /* QDFact = Rate/320 under -r (p64.c:576)                       */
  int32_t qoffs;           /* QOffs  = 1 (p64.c:577)                                       */
  int32_t reserved[2];
} p64b_rate_control;
int p64b_ctx_set_rate_control(p64b_ctx *ctx, const p64b_rate_control *rc);

/* Rate-control split (-r / -x): the quantiser of GOB g is chosen by the host from the bits written so
 * far (ExecuteQuantization, p64.c:458-481, 697-702), so quantise..reconstruct runs per GOB.
 *   begin : upload + motion estimation for the whole frame (quantiser independent)
 *   gob   : decision..reconstruct for GOB `gob` of every stream with per-stream GQUANT quant[n_streams];
 *           outputs cover that GOB only: mbs [n_streams][33], levels [n_streams][33][6][64]
 *   end   : MBs the host overrode to "type 4, zero vector" on buffer overflow (p64.c:776-783) are
 *           re-reconstructed as a copy and their LastIntra counter fixed; then SwapFS.
 *           overflow [n_streams][num_mb] (non-zero = overridden) or NULL. */
int p64b_ctx_frame_begin(p64b_ctx *ctx, const p64b_step *step, const uint8_t *src);
int p64b_ctx_encode_gob(p64b_ctx *ctx, const p64b_step *step, int gob, const uint8_t *quant, p64b_mb *mbs,
                        int8_t *levels);
int p64b_ctx_frame_end(p64b_ctx *ctx, const uint8_t *overflow);

/* Decoder's inverse half (DecompressMDU p64.c:1179-1237 + DecodeSaveMDU p64.c:971-1013) for one picture of
 * every stream: inverse quantise, Chen IDCT, prediction (overlay / MC / half-vector chroma / loop filter), clamp, store.
 *   mbs    [n_streams][num_mb] GOB-major: mtype, cbp, mvx/mvy (full vectors, MVD prediction already undone), quant = UseQuant;
 *          bit 0 of `reserved` set iff the stream transmitted the macroblock -- others keep the previous picture's samples
 *   levels [n_streams][num_mb][6][64] transmission order, as the VLC decoder delivers them (intra DC as uint8)
 * The decoded picture becomes the reference store: p64b_ctx_download_recon() reads it. */
int p64b_ctx_decode_frames(p64b_ctx *ctx, const p64b_mb *mbs, const int8_t *levels);

/* Motion estimation alone on device-resident luma planes (config 4 microbenchmark):
 * ref/cur [n_pairs][W*H], out [n_pairs][num_mb raster]. n_pairs may exceed the context's stream count. */
int p64b_ctx_motion_estimation_dev(p64b_ctx *ctx, const uint8_t *ref_dev, const uint8_t *cur_dev, int n_pairs,
                                   int me_mode, int search_limit, p64b_me *out_dev);

/* Test hook: the three-step search (out_dev = its records) that additionally writes every macroblock's 31x31 SAD
 * surface, surface_dev uint32 [n_pairs][num_mb][31][31] indexed [dy+15][dx+15]; 0xffffffff marks positions the
 * reference's legality rule excludes (me.c:292-293); [15][15] always holds SAD(0,0) (me.c:262-271). */
int p64b_ctx_sad_surface_dev(p64b_ctx *ctx, const uint8_t *ref_dev, const uint8_t *cur_dev, int n_pairs,
                             p64b_me *out_dev, uint32_t *surface_dev);

/* The reference's ME arrays for one stream, raster MB order (for -o programs and tests; me.c:49-59). */
int p64b_ctx_me_records(p64b_ctx *ctx, int stream, p64b_me *out);
/* The current reference frame store (= last reconstructed frame, CFS after SwapFS), for -l statistics
 * (stat.c:52) and closed-loop tests. */
int p64b_ctx_download_recon(p64b_ctx *ctx, int stream, uint8_t *yuv);
/* Frame statistics (Statistics / StatisticsMem, stat.c:52-130, printed by -l): the exact integer sums of
 * the last coded frame of every stream, source frame vs its reconstruction (CFS after SwapFS), per plane Y, Cb, Cr --
 * accumulated on the device; p64b_stat_from_sums() turns one record into the doubles the reference prints. */
typedef struct p64b_plane_stats {
  uint64_t n;            /* samples: width*height of the plane (top)                      */
  uint64_t sum_src;      /* sum of the source samples             (rvalue,   stat.c:97)  */
  uint64_t sum_rec;      /* sum of the reconstructed samples      (value,    stat.c:98)  */
  uint64_t sum_sq_err;   /* sum of (reconstruction - source)^2    (squared,  stat.c:99)  */
  uint64_t sum_sq_src;   /* sum of source^2                       (rsquared, stat.c:100) */
  uint32_t hist[256];    /* histogram of the reconstruction       (Values,   stat.c:102) */
} p64b_plane_stats;
typedef struct p64b_stat { double mean, mse, snr, mrsnr, psnr, entropy; } p64b_stat;   /* STAT, globals.h */
int p64b_ctx_statistics(p64b_ctx *ctx, p64b_plane_stats *out /* [n_streams][3] */);
void p64b_stat_from_sums(const p64b_plane_stats *sums, p64b_stat *out);
/* LastIntra counters [num_mb] GOB-major (p64.c:213). */
int p64b_ctx_last_intra(p64b_ctx *ctx, int stream, uint8_t *out);
/* number of kernel launches issued by this context so far */
int64_t p64b_ctx_launches(const p64b_ctx *ctx);
/* Steps of p64b_ctx_wait_bits whose frame was larger than the download budget (completed by a second copy). */
int64_t p64b_ctx_second_copies(const p64b_ctx *ctx);
/* Packed 4-byte SAD operations (VABSDIFF4.U8.ACC lane-instructions) the motion-estimation kernel has EXECUTED so far in
 * its search sweeps, counted on the device; reset != 0 zeroes the counter.  The exhaustive search leaves a pass early when
 * no candidate of the pass can still win (the warp-wide form of ComputeError's early exit, me.c:122-170), so this is what
 * the issue-rate roofline is measured with; the algorithmic count (every legal candidate in full) is content independent. */
int p64b_ctx_me_executed(p64b_ctx *ctx, uint64_t *packed_sad_ops, int reset);
/* Per-kernel device timing (CUDA events on the launching stream around every launch) for the roofline report.
 * profile(ctx,1) clears and starts recording, profile(ctx,0) stops; profile_read sums the recorded launches:
 * ms_total[3], count[3] -- index 0 = motion-estimation kernel, 1 = macroblock (DCT/quant/recon) kernel,
 * 2 = the entropy-coding kernels of p64b_ctx_submit_bits (one interval around the three of them). */
int p64b_ctx_profile(p64b_ctx *ctx, int enable);
int p64b_ctx_profile_read(p64b_ctx *ctx, double *ms_total, int32_t *count);
/* Measured issue peak of the packed 4-byte SAD instruction (VABSDIFF4.U8.ACC) on this device, in
 * packed ops per second: the denominator of the ME roofline (no datasheet figure exists). */
int p64b_measure_sad_peak(int device, double *ops_per_s, double *sm_clock_mhz);
/* Measured host-to-device copy rate (GB/s) of `reps` back-to-back copies of `bytes` from `host` (pinned): the ceiling of
 * the upload-bound end-to-end path. */
int p64b_measure_h2d(int device, const void *host, size_t bytes, int reps, double *gb_per_s);
/* Host link probe for the end-to-end attribution.  `reps` uploads of up_bytes each, taken round robin
 * from the up_sets buffers up[0..up_sets-1] (distinct sets larger than the host's last-level cache behave like the
 * end-to-end path; one set repeated does not), and/or `reps` downloads of down_bytes into `down`, on two copy streams.
 * mode: 1 = uploads only, 2 = downloads only, 3 = both directions at the same time.  Rates in GB/s (0 where not run). */
int p64b_measure_link(int device, const void *const *up, int up_sets, size_t up_bytes, void *down, size_t down_bytes, int reps,
                      int mode, double *up_gb_per_s, double *down_gb_per_s);
/* Upload rate (GB/s) of every listed device while ALL of them upload at the same time (one host thread per device, 8 x 32 MB
 * from pinned memory): the GPUs of a box share host uplinks, not always evenly.  p64b_enc_params.balance_links partitions by it. */
int p64b_probe_links(const int32_t *devices, int n_devices, double *gb_per_s);
/* Debug builds (-DP64B_BOUNDS_CHECK) check every shared-memory access of the motion-estimation and macroblock kernels against
 * the region its thread may touch (compute-sanitizer is not available on every pool).  Returns 1 and the number of violations
 * on `device` since the library was loaded (+ the kernels.cuh line of the last one) from such a build, 0 from the product build. */
int p64b_debug_oob(int device, uint32_t *violations, uint32_t *last_line);
/* Self-test of the sequence encoder's host thread pool (no GPU needed): `rounds` parallel loops of `items` items on `workers`
 * threads + the caller; 0 if every item ran exactly once in every round. */
int p64b_debug_pool_selftest(int workers, int items, int rounds);
/* p64b_host_alloc with flags: 1 = write-combined (cudaHostAllocWriteCombined). */
void *p64b_host_alloc_flags(size_t bytes, int flags);

/* ---------------------------------------------------------------------------------------------
 * (2) host bit stream: headers + VLC (marker.c, codec.c, huffman.c, stream.c)
 * ------------------------------------------------------------------------------------------- */
typedef struct p64b_bits p64b_bits;

p64b_bits *p64b_bits_create(int image_type);
void p64b_bits_destroy(p64b_bits *b);
/* WritePictureHeader, marker.c:103-137 (PSC, TR, PTYPE, PSPARE for NTSC: p64.c:408-423). */
void p64b_bits_picture_header(p64b_bits *b, int temporal_reference);
/* WriteGOBHeader, marker.c:182-209; gob is CurrentGOB (0-based); QCIF numbers GOBs 1,3,5 (p64.c:709-711). */
void p64b_bits_gob_header(p64b_bits *b, int gob, int gquant);
/* WriteMDU's write half: WriteMBHeader (marker.c:288-354) + EncodeDC/EncodeAC/CBPEncodeAC (codec.c:96-205,
 * 346-355) for MB `mdu` (0..32) of the current GOB. */
void p64b_bits_mb(p64b_bits *b, int mdu, const p64b_mb *rec, const int8_t *levels);
/* mputv, stream.c:193-205: nbits (<= 32) raw bits, MSB first. */
void p64b_bits_put(p64b_bits *b, uint32_t value, int nbits);
/* mwtell, stream.c:233-238: bits written so far. */
int64_t p64b_bits_tell(const p64b_bits *b);
/* mwclose, stream.c:142-152: pad the last byte with 1-bits. Returns the byte count. */
size_t p64b_bits_finish(p64b_bits *b);
const uint8_t *p64b_bits_data(const p64b_bits *b, size_t *nbytes);
/* The per-frame counters PrintFrameStatistics() reports (p64.c:198-211, 640-649, 1299-1332), accumulated by
 * p64b_bits_mb() since the last p64b_bits_counters_reset(). */
typedef struct p64b_frame_counters {
  int32_t mb_attribute_bits;    /* MacroAttributeBits, marker.c:354 */
  int32_t mv_bits;              /* MotionVectorBits,   marker.c:344 */
  int32_t eob_bits;             /* EOBBits,            codec.c:129, 204 */
  int32_t y_bits, u_bits, v_bits;   /* Y/U/VCoefBits,  p64.c:949-951 */
  int32_t number_nz;            /* NumberNZ,           codec.c:125, 169, 200, 351 */
  int32_t q_sum, q_use;         /* QSum, QUse,         p64.c:804-805 */
  int32_t macro_type_freq[10], y_type_freq[10], uv_type_freq[10];   /* p64.c:807, 937-938 */
  int32_t total_bits, last_bits;            /* TotalBits, LastBits (p64.c:654-658); filled by p64b_enc_frame_counters */
  int32_t buffer_contents, buffer_size;     /* BufferContents(), BufferSize() as printed under -r (p64.c:1304-1308) */
} p64b_frame_counters;
void p64b_bits_counters(const p64b_bits *b, p64b_frame_counters *out);
void p64b_bits_counters_reset(p64b_bits *b);
/* empties the writer (bytes, predictors, counters) */
void p64b_bits_reset(p64b_bits *b);

/* ---------------------------------------------------------------------------------------------
 * (2b) YUV4MPEG2 reader of the ingest path (vidinput.c / y4m_input.c:556-768): parses the header and FRAME markers
 * with the reference reader's rules and reads each frame's payload, unconverted, into the caller's staging buffer.
 * ------------------------------------------------------------------------------------------- */
typedef struct p64b_y4m p64b_y4m;
typedef struct p64b_y4m_info {
  int32_t width, height;      /* W, H                                                   */
  int32_t fps_n, fps_d;       /* F                                                      */
  int32_t par_n, par_d;       /* A (0:0 when absent)                                    */
  int32_t chroma;             /* P64B_CHROMA_* from the C tag ("420" when absent)       */
  int32_t interlace;          /* I tag character, '?' when absent                       */
  int64_t frame_bytes;        /* payload bytes per frame in the file                    */
} p64b_y4m_info;
int p64b_y4m_open(p64b_y4m **out, const char *path /* "-" = stdin */);
void p64b_y4m_close(p64b_y4m *y);
int p64b_y4m_get_info(const p64b_y4m *y, p64b_y4m_info *info);
/* 1 = a frame was read into dst[frame_bytes], 0 = end of file, < 0 = error */
int p64b_y4m_read_frame(p64b_y4m *y, uint8_t *dst);
int64_t p64b_y4m_payload_bytes(int width, int height, int chroma);

/* ---------------------------------------------------------------------------------------------
 * (2c) decoder: the sequential bit-stream parser (ReadPictureHeader / ReadGOBHeader / ReadMBHeader marker.c:144-450,
 * DecodeDC / DecodeAC / CBPDecodeAC codec.c:214-340) on the host + the inverse half on the device, driven like
 * p64DecodeSequence (p64.c:1022-1126).  One stream per decoder object.
 * ------------------------------------------------------------------------------------------- */
typedef struct p64b_dec p64b_dec;
/* `data` must stay valid for the decoder's life time.  The image type comes from the first picture header (p64.c:1071-1081). */
int p64b_dec_create(p64b_dec **out, int device, const uint8_t *data, size_t nbytes);
void p64b_dec_destroy(p64b_dec *d);
int p64b_dec_image_type(const p64b_dec *d);
/* Decodes the next picture into yuv[frame_bytes].  Returns 1 with *repeat = how many times p64DecodeSequence writes it
 * (temporal-reference gaps repeat a picture, p64.c:1047-1054), 0 at the end of the stream, < 0 on error. */
int p64b_dec_next_picture(p64b_dec *d, uint8_t *yuv, int *repeat);
/* The parser alone (no device): fills mbs[num_mb] / levels[num_mb][6][64] for the next picture as p64b_ctx_decode_frames
 * wants them; *temporal_reference = TR of its picture header.  Returns 1 / 0 / < 0 like p64b_dec_next_picture. */
typedef struct p64b_parser p64b_parser;
int p64b_parser_create(p64b_parser **out, const uint8_t *data, size_t nbytes);
void p64b_parser_destroy(p64b_parser *p);
int p64b_parser_image_type(const p64b_parser *p);
int p64b_parser_next_picture(p64b_parser *p, p64b_mb *mbs, int8_t *levels, int *temporal_reference, int *repeat);

/* ---------------------------------------------------------------------------------------------
 * (3) sequence driver: p64EncodeSequence / p64EncodeFrame / p64EncodeGOB (p64.c:524-786) for a batch
 * ------------------------------------------------------------------------------------------- */
typedef struct p64b_enc p64b_enc;

#define P64B_MAX_DEVICES 16
typedef struct p64b_enc_params {
  int32_t image_type;
  int32_t n_streams;
  int32_t device;
  int32_t start_frame;     /* -a : only enters TemporalReference (p64.c:637)                       */
  int32_t initial_quant;   /* -q ; 0 = default (8, or 10000000/Rate under -r: p64.c:574-590)       */
  int32_t rate;            /* -r bits/s; 0 = fixed quantiser                                       */
  int32_t frame_rate;      /* -f numerator (default 30000)                                         */
  int32_t frame_rate_div;  /* -f denominator (default 1001)                                        */
  int32_t frame_skip;      /* -k (default 1)                                                       */
  int32_t me_mode;         /* P64B_ME_TSS (stock) | P64B_ME_FULL                                   */
  int32_t search_limit;    /* -i (default 15)                                                      */
  int32_t force_intra;     /* `-o < test.intra`                                                    */
  int32_t vlc_threads;     /* host threads for the per-stream VLC (0 = one per core, capped)       */
  int32_t host_vlc;        /* 1: entropy-code (and, under -r, run the rate control) on the host through the per-GOB
                              calls; default 0: both on the device (p64b_ctx_submit_bits)                       */
  int32_t input_chroma;    /* P64B_CHROMA_*: layout of the frames given to p64b_enc_encode (default 420jpeg)    */
  int32_t last_frame;      /* -b LastFrame + 1, or 0 = unknown (then: the last frame actually coded).  Only the trailing
                              picture header needs it: its TR is min(CurrentFrame, LastFrame+1) % 32 (p64.c:600-602), which
                              with -k > 1 depends on where -b stops between two coded frames                    */
  /* Multi-GPU: n_devices > 0 partitions the streams over devices[0..n_devices-1] in contiguous blocks whose
   * sizes differ by at most one (the same split as p64_b200/shard.py); every device gets its own p64b_ctx and its own
   * host worker thread, there is no exchange between them (streams are independent; one stream is sequential across
   * frames, p64.c:661).  n_devices = 0: the single `device` above.  A device may be listed more than once. */
  int32_t n_devices;
  int32_t devices[P64B_MAX_DEVICES];
  /* 1: size the blocks in proportion to the upload rate each device gets while all of them upload (p64b_probe_links,
   * measured at creation) instead of equally -- on a box whose GPUs share host uplinks unevenly the equal split waits for
   * the slowest link (measured at 8 GPUs: +15 % end to end).  The bytes of a stream never depend on the partition. */
  int32_t balance_links;
} p64b_enc_params;

void p64b_enc_default_params(p64b_enc_params *p);
int p64b_enc_create(p64b_enc **out, const p64b_enc_params *p);
void p64b_enc_destroy(p64b_enc *e);
/* Encode the next frame of every stream: src [n_streams][frame_bytes] (host); with input_chroma set,
 * [n_streams][p64b_raw_frame_bytes()] unconverted Y4M payloads.  `src` is consumed before the call returns (copied
 * into pinned staging), so the caller may reuse it at once.  p64b_enc_staging() is the encoder's own pinned upload buffer
 * for the NEXT frame: a reader may fill it in place and pass it as `src` (no extra copy); it changes after every encode.
 * With the device-side coder (host_vlc = 0) frames are pipelined three deep: the call enqueues this frame and collects
 * the one submitted three calls earlier, p64b_enc_finish() collects the rest -- p64b_enc_data() / p64b_enc_overflows() /
 * p64b_enc_first_frame_bits() are complete after p64b_enc_finish(). */
int p64b_enc_encode(p64b_enc *e, const uint8_t *src);
uint8_t *p64b_enc_staging(p64b_enc *e);
/* Trailing picture header + padding (p64.c:600-605) for every stream. Call once after the last frame. */
int p64b_enc_finish(p64b_enc *e);
/* The stream's bytes so far (complete after p64b_enc_finish). */
const uint8_t *p64b_enc_data(const p64b_enc *e, int stream, size_t *nbytes);
p64b_ctx *p64b_enc_ctx(p64b_enc *e);     /* multi-device encoders: the context of the first device */
/* Number of device partitions (1 for a single-device encoder) and the first stream / stream count of partition k. */
int p64b_enc_partitions(const p64b_enc *e);
int p64b_enc_partition(const p64b_enc *e, int k, int *device, int *first_stream, int *n_streams);
/* Statistics the reference prints: buffer overflows (p64.c:779), bits of the first frame (p64.c:668). */
int64_t p64b_enc_overflows(const p64b_enc *e, int stream);
int64_t p64b_enc_first_frame_bits(const p64b_enc *e, int stream);
/* PrintFrameStatistics()'s counters for the frame just coded (-l).  Needs host_vlc = 1 (the categories are counted by
 * the host bit writer); P64B_EINVAL otherwise. */
int p64b_enc_frame_counters(const p64b_enc *e, int stream, p64b_frame_counters *out);

#ifdef __cplusplus
}
#endif
#endif /* P64_B200_H */
