// sm_100a kernels of the H.261 hot path.  All arithmetic is integer except the MTYPE decision (double),
// exactly as in the reference; citations "file:line" are into maikmerten/p64.
//
//   me_search_kernel    block matching over the legal part of the 31x31 SAD surface: exhaustive argmin (with the exact
//                       warp-wide early exit), the three-step walk, + the VAR/VAROR/MWOR statistics   (me.c:187-363)
//   mb_encode_kernel    MTYPE decision, prediction (MC / half-vector chroma / loop filter), residual,
//                       Chen DCT, quantise, zig-zag, CBP + type-4/7 fallback, inverse quantise,
//                       Chen IDCT, reconstruct                        (p64.c:734-773, 823-913, 935-1013)
//   overflow_patch_kernel  MBs overridden to "type 4, zero vector" on buffer overflow (p64.c:776-783)
//   mb_decode_kernel    the decoder's inverse half                    (p64.c:971-1013, 1179-1237)
//   plane_stats_kernel  per-plane sums and histogram for the -l statistics  (stat.c:73-130)
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/p64_b200.h"

namespace p64b {

struct Geom {
  int W, H;        // luma size
  int mbw, mbh;    // macroblocks per row / column
  int ngob, nmb;   // GOBs, macroblocks per frame
  int qcif;        // GOB layout selector (io.c:730-741)
  int frame_bytes; // W*H*3/2
};

// ---------------------------------------------------------------------------------------------------
// Motion estimation
// ---------------------------------------------------------------------------------------------------
// Persistent kernel, ONE WARP PER MACROBLOCK: every warp owns a slice of shared memory and pulls macroblocks from a
// global work queue (an atomic counter; the index of the macroblock after next is requested one macroblock ahead,
// so its latency is hidden) with no CTA-wide synchronisation at all.  Per macroblock:
//   * the 47x48-byte search window and the 16x16 current block arrive by TMA tile loads (out-of-frame samples
//     zero-filled by the TMA unit) into a double buffer: the loads for the warp's NEXT macroblock are issued before
//     the current one is processed, so the TMA latency is never exposed;
//   * the window is replicated shifted by 1..3 bytes (TMA needs a 16-byte aligned innermost start), so that every
//     packed-SAD operand in the hot loop is an aligned 32-bit shared-memory word (no PRMT/funnel shift in the loop);
//   * lane = dx position; the warp sweeps the legal dy positions in chunks of 10 candidates per thread: a
//     thread keeps the whole current block in 64 registers and slides down the window rows (4 LDS.32 per row),
//     feeding 10 independent VABSDIFF4.U8.ACC chains; the exhaustive search eliminates candidates exactly, warp-wide
//     first and then one by one (see full_chunk below).
// Only the part of the 31x31 surface that the reference can ever look at is evaluated: the legal positions
// (me.c:212-213, 292-293) inside the search range -- 30x30 for FastBME -i 31 inside the frame ([-15,14],
// me.c:206-208), 31x31 for StepBME, 15 or 16 wide/high for macroblocks on a frame edge.  With <= 16 legal dx the
// two half-warps take the two halves of the dy range, so edge macroblocks cost 1/2 and corner macroblocks about 1/3
// of an interior one.  SAD(0,0), which both searches probe first whatever the legality (me.c:203, 271), is computed
// separately (two packed SADs per lane).
// Shared-memory banks: the word of copy k = (dx+16)&3 at word offset q = (dx+16)>>2 lies in bank 8k+q+const
// (copy 0 starts 128-byte aligned, copy k at 584(k-1)+8 words from an aligned base): conflict-free for the 31 dx of a
// warp; in the two-half mapping the second half is 15 rows = 180 words further: +20 banks -> the complementary banks.
constexpr int ME_WARPS = 4;            // warps per CTA (independent workers)
constexpr int ME_THREADS = 32 * ME_WARPS;
constexpr int ME_WIN_ROWS = 47;        // window rows y0-15 .. y0+31
constexpr int ME_ROW_WORDS = 12;       // 48 bytes: x0-16+k .. x0+31+k
constexpr int ME_WIN_BYTES = ME_WIN_ROWS * 48;
// per-warp shared memory, in 32-bit words
constexpr int ME_BUF_WORDS = 608 + 64;                 // TMA window (564 used, padded to 128-byte multiple) + current block
constexpr int ME_SHIFT_OFF = 2 * ME_BUF_WORDS;         // byte-shifted copies 1..3: copy k at ME_SHIFT_OFF + 8 + 584 (k-1)
constexpr int ME_COPY_WORDS = 584;
constexpr int ME_SAD_OFF = ME_SHIFT_OFF + 8 + 3 * ME_COPY_WORDS + 8;      // [31][31] surface (ME_V_SURF only); survivor list (ME_V_FULL)
#ifndef P64B_ME_RA
#define P64B_ME_RA 8
#endif
#ifndef P64B_ME_DENSE_T
#define P64B_ME_DENSE_T 48
#endif
constexpr int ME_RA = P64B_ME_RA;               // block rows every candidate of a surviving chunk accumulates before the per-candidate elimination
constexpr int ME_DENSE_T = P64B_ME_DENSE_T;     // more survivors than this in one chunk (of <= 320): the chunk continues densely (sliding window)
constexpr int ME_LIST_WORDS = 160;              // survivor list: flushed when a further chunk might not fit
constexpr int ME_WARP_WORDS_FULL = (ME_SAD_OFF + ME_LIST_WORDS + 31) / 32 * 32;
constexpr int ME_WARP_WORDS_SURF = (ME_SAD_OFF + 31 * 31 + 31) / 32 * 32;
constexpr int ME_CTA_WORDS = 3 * 32 /*pen tables*/ + 2 * 2 * ME_WARPS /*mbarriers*/ + 16;
constexpr int ME_SMEM_FULL = 128 + 4 * (ME_WARPS * ME_WARP_WORDS_FULL + ME_CTA_WORDS);
constexpr int ME_SMEM_SURF = 128 + 4 * (ME_WARPS * ME_WARP_WORDS_SURF + ME_CTA_WORDS);
constexpr uint32_t ME_ILLEGAL = 0x10000u;   // start value of an illegal candidate's accumulator: above any SAD (<= 65280)

// CTA-uniform legal ranges in surface positions d+15, per macroblock-column / row class
// (0 = first, 1 = interior, 2 = last), computed on the host: one byte per class (hi stored +1, 0 = empty).
struct MeRanges { uint32_t xlo, xhi, ylo, yhi; };
struct MeArgs {
  MeRanges rg;
  int me_mode;
  int mbw, mbh, n_pairs;
  uint32_t magic_pp, shift_pp, magic_w;   // n / (mbw*mbh) = umulhi(n, magic_pp) >> shift_pp;  r / mbw = r * magic_w >> 16
  uint32_t* queue;         // [2] work counters: this launch takes macroblocks from queue[parity] and zeroes the other
  int parity;
  uint32_t m8, m16, m24, m2048;   // 1<<8, 1<<16, 1<<24, 2048 in registers: keeps those multiplies on the FMA pipe
  p64b_me* out;
  uint32_t* surface;
  unsigned long long* executed;   // += candidate rows accumulated per lane, summed over warps (nullptr: not counted)
};

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // VABSDIFF4.U8.ACC
  return d;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- debug build (-DP64B_BOUNDS_CHECK): every shared-memory access of me_search_kernel and mb_encode_kernel is checked against the
// region its thread may touch (the warp's slice + the CTA tables; the thread's tile + the CTA exchange array).  compute-sanitizer
// is closed on this pool, so this is how the hand-computed offsets of the warp-synchronous kernels are validated
// (tests/test_bounds_check.py runs edge / corner macroblocks and ragged stream counts through such a build).  Violations are
// counted in g_oob[0] (g_oob[1] = source line of the last one) and the access is redirected to the region's first word.
__device__ unsigned int g_oob[2];
#ifdef P64B_BOUNDS_CHECK
__device__ __forceinline__ uint32_t smck_fix(uint32_t ad, uint32_t bytes, uint32_t lo, uint32_t hi, uint32_t lo2, uint32_t hi2, int line) {
  if ((ad >= lo && ad + bytes <= hi) || (ad >= lo2 && ad + bytes <= hi2)) return ad;
  atomicAdd(&g_oob[0], 1u);
  g_oob[1] = (unsigned)line;
  return lo;
}
#define P64B_SM_T(T, p, lim) (*reinterpret_cast<T*>(__cvta_shared_to_generic(smck_fix(smem_u32(p), (uint32_t)sizeof(T), (lim).lo, (lim).hi, (lim).lo2, (lim).hi2, __LINE__))))
#else
#define P64B_SM_T(T, p, lim) (*reinterpret_cast<T*>(p))
#endif
struct SmLim { uint32_t lo, hi, lo2, hi2; };        // two [lo, hi) windows of shared-memory addresses (unused in the product build)
#define SMR(p) P64B_SM_T(const uint32_t, (p), lim)
#define SMR2(p) P64B_SM_T(const uint2, (p), lim)
#define SMR4(p) P64B_SM_T(const uint4, (p), lim)
#define SMW(p) P64B_SM_T(uint32_t, (p), lim)
#define SMW4(p) P64B_SM_T(uint4, (p), lim)
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA: 3-D tiled bulk tensor load global -> shared, completion on an mbarrier (UTMALDG in SASS)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// next position of the work queue, by one lane.  Inline PTX: the compiler turns a plain atomicAdd into its warp-aggregated
// form, whose result broadcast (SHFL) waits for the atomic at once -- 3 % of the kernel's warp time in the ncu source view --
// while the value is only needed one macroblock later.
__device__ __forceinline__ uint32_t queue_take(uint32_t* counter, int lane) {
  uint32_t t = 0;
  // every lane presents its own address (lane 0 the counter, the others are predicated off): nothing to aggregate
  asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %2, 0;\n@p atom.global.add.u32 %0, [%1], 1;\n}" : "+r"(t) : "l"(counter + lane), "r"(lane) : "memory");
  return t;
}

constexpr int ME_V_FULL = 0, ME_V_SURF = 1, ME_V_TSS = 2;

#ifndef P64B_ME_PRUNE_R1
#define P64B_ME_PRUNE_R1 4
#endif
constexpr int ME_PRUNE_R1 = P64B_ME_PRUNE_R1;   // first warp-wide check point of a chunk (block rows done); the second is ME_RA

// block rows [R0, R1) of all NC candidates: window rows R0 .. R1-1+NC-1
template <int NC, int R0, int R1>
__device__ __forceinline__ void sweep_rows(const uint32_t* __restrict__ base, const uint32_t (&c)[16][4], uint32_t (&a)[NC], const SmLim& lim) {
#pragma unroll
  for (int t = R0; t < R1 - 1 + NC; t++) {
    const uint32_t r0 = SMR(base + t * ME_ROW_WORDS + 0), r1 = SMR(base + t * ME_ROW_WORDS + 1);
    const uint32_t r2 = SMR(base + t * ME_ROW_WORDS + 2), r3 = SMR(base + t * ME_ROW_WORDS + 3);
#pragma unroll
    for (int j = 0; j < NC; j++) {
      const int i = t - j;
      if (i >= R0 && i < R1) {
        a[j] = sad4(r0, c[i][0], a[j]); a[j] = sad4(r1, c[i][1], a[j]);
        a[j] = sad4(r2, c[i][2], a[j]); a[j] = sad4(r3, c[i][3], a[j]);
      }
    }
  }
}
template <int NC>
__device__ __forceinline__ bool sweep_hopeless(const uint32_t (&a)[NC], bool xok, uint32_t bound) {
  uint32_t m = a[0];
#pragma unroll
  for (int j = 1; j < NC; j++) m = min(m, a[j]);
  return __reduce_min_sync(0xffffffffu, xok ? m : 0xffffffffu) > bound;
}

// Surface variant (test hook): one pass of NC candidates per thread at dy positions yb .. yb+NC-1 in full (accumulators start at
// 0 or ME_ILLEGAL from the row table); the legal entries go into the surface in shared memory.
template <int VARIANT, int NC>
__device__ __forceinline__ uint32_t sweep_pass(const uint32_t* __restrict__ colbase, const uint32_t (&c)[16][4],
                                               const uint32_t* pen, uint32_t* s_sad, int xi, int yb, bool xok,
                                               uint32_t m2048, uint32_t& units, const SmLim& lim) {
  const uint32_t* base = colbase + yb * ME_ROW_WORDS;
  uint32_t a[NC];
#pragma unroll
  for (int j = 0; j < NC; j++) a[j] = SMR(pen + yb + j);
  sweep_rows<NC, 0, 16>(base, c, a, lim);
  units += (uint32_t)(NC * 16);
  if (VARIANT == ME_V_SURF) {
#pragma unroll
    for (int j = 0; j < NC; j++)
      if (xok && a[j] < ME_ILLEGAL) SMW(s_sad + (yb + j) * 31 + xi) = a[j];     // the rest keeps the 0xffffffff prefill
  }
  uint32_t best = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < NC; j++) best = min(best, a[j] * m2048 + (uint32_t)j);      // IMAD (FMA pipe), not LEA (ALU pipe)
  return best + (uint32_t)yb;
}

// ---- exhaustive search: exact elimination, warp-wide first, then per candidate (ME_V_FULL) ---------------------------------
// The reference's ComputeError leaves a candidate as soon as its partial sum reaches the best SAD so far (me.c:65, 122, 133,
// 170); that never changes a decision because acceptance needs a strictly smaller FULL sum (me.c:220).  Here every candidate
// whose partial sum is strictly above `bound` -- the full SAD of some candidate already evaluated, SAD(0,0) at first -- is
// dropped: it can neither win nor tie.  Everything else is evaluated in full and enters the minimum of the key
// SAD << 11 | scan order, so the winner is the reference's for any content and any evaluation order.
//   chunk = 10 dy x 32 dx, accumulators in registers, sliding window:
//     A. block rows 0..R1-1 of all candidates; the chunk is dropped when EVERY candidate is above the bound; rows R1..RA-1; again;
//     B. (only chunks that survive -- typically the one that holds the match) the candidate with the chunk's smallest
//        partial sum is evaluated in full by the whole warp (2 packed SADs per lane) and becomes the bound;
//     C. survivors (partial <= bound) are found with one ballot per dy row.  More than ME_DENSE_T: the chunk continues densely
//        (rows RA..15); else they are appended to a list in shared memory (partial | dy | dx);
//   list: 32 survivors at a time, ONE PER LANE with two accumulator chains, finish their rows RA..15 from the byte-shifted
//     window copies (aligned words), leaving half way when no lane is still at or below the bound.
// What this buys depends on the content (tools/me_prune_sim.py models it): on the bench clip the search executes 0.53 of the
// algorithmic SAD operations instead of 0.70 with the warp-wide exit alone.  Variants measured on a B200 and not kept:
// elimination after 6 rows with every survivor finished from the list (0.48 executed, but 2400 instead of 2240 instructions
// per macroblock: the bookkeeping of three compactions per macroblock costs what the SADs save), and a fully unrolled 30-row
// column (code beyond the 32 KB instruction cache: 1.5x slower).
// full SAD of the candidate at surface position (cx, cy), cooperatively: lane -> (block row lane/2, half lane%2)
__device__ __forceinline__ uint32_t coop_sad(const uint32_t* win, const uint32_t* s_cur, int cx, int cy, int lane, const SmLim& lim) {
  const int i = lane >> 1, wc = (lane & 1) * 2, oo = cx + 1, sh = (oo & 3) * 8;
  const uint32_t* rp = win + (cy + i) * ME_ROW_WORDS + (oo >> 2) + wc;
  const uint32_t w0 = SMR(rp), w1 = SMR(rp + 1), w2 = SMR(rp + 2);
  const uint32_t ra = __funnelshift_r(w0, w1, sh), rb = __funnelshift_r(w1, w2, sh);
  const uint2 cw = SMR2(s_cur + i * 4 + wc);
  return __reduce_add_sync(0xffffffffu, sad4(rb, cw.y, sad4(ra, cw.x, 0u)));
}

// The list's survivors, one per lane: rows RA..15 on two accumulator chains, one exit half way.
__device__ __forceinline__ void sparse_finish(const uint32_t* win, const uint32_t* shifted, const uint32_t (&c)[16][4], const uint32_t* list,
                                              int nlist, int lane, uint32_t m2048, uint32_t& bound, uint32_t& best, uint32_t& units, const SmLim& lim) {
  constexpr int S1 = (ME_RA + 16) / 2;
#pragma unroll 1
  for (int i0 = 0; i0 < nlist; i0 += 32) {
    const bool ok = i0 + lane < nlist;
    const uint32_t e = ok ? SMR(list + i0 + lane) : 0u;
    uint32_t p = ok ? (e & 0xffffu) : ME_ILLEGAL, q = 0;
    if (!__any_sync(0xffffffffu, p <= bound)) continue;              // the bound has moved since they were listed
    const int x = (e >> 21) & 31, y = (e >> 16) & 31, o = x + 1, k = o & 3;
    const uint32_t* b = (k ? shifted + (k - 1) * ME_COPY_WORDS : win) + (o >> 2) + y * ME_ROW_WORDS;
#pragma unroll
    for (int r = ME_RA; r < S1; r++) {
      p = sad4(SMR(b + r * ME_ROW_WORDS + 0), c[r][0], p); q = sad4(SMR(b + r * ME_ROW_WORDS + 1), c[r][1], q);
      p = sad4(SMR(b + r * ME_ROW_WORDS + 2), c[r][2], p); q = sad4(SMR(b + r * ME_ROW_WORDS + 3), c[r][3], q);
    }
    units += S1 - ME_RA;
    if (!__any_sync(0xffffffffu, p + q <= bound)) continue;
#pragma unroll
    for (int r = S1; r < 16; r++) {
      p = sad4(SMR(b + r * ME_ROW_WORDS + 0), c[r][0], p); q = sad4(SMR(b + r * ME_ROW_WORDS + 1), c[r][1], q);
      p = sad4(SMR(b + r * ME_ROW_WORDS + 2), c[r][2], p); q = sad4(SMR(b + r * ME_ROW_WORDS + 3), c[r][3], q);
    }
    units += 16 - S1;
    p += q;                                   // (lanes without a survivor carry ME_ILLEGAL: never the minimum)
    best = min(best, p * m2048 + (uint32_t)(1 + x * 32 + y));
    bound = min(bound, __reduce_min_sync(0xffffffffu, p));
  }
}

// One chunk of 10 dy rows at surface rows yb .. yb+9 of this lane's column xi.  Returns the list length.
__device__ __forceinline__ int full_chunk(const uint32_t* colbase, const uint32_t* win, const uint32_t* s_cur, const uint32_t (&c)[16][4],
                                          const uint32_t* pen, uint32_t* list, int nlist, int xi, int yb, bool xok, int lane, uint32_t m2048,
                                          uint32_t& bound, uint32_t& best, uint32_t& units, bool& dense_mode, const SmLim& lim) {
  constexpr int NC = 10, R1 = ME_PRUNE_R1;
  const uint32_t* base = colbase + yb * ME_ROW_WORDS;
  uint32_t a[NC];
#pragma unroll
  for (int j = 0; j < NC; j++) a[j] = SMR(pen + yb + j);
  sweep_rows<NC, 0, R1>(base, c, a, lim);
  units += (uint32_t)(NC * R1);
  if (sweep_hopeless<NC>(a, xok, bound)) return nlist;
  sweep_rows<NC, R1, ME_RA>(base, c, a, lim);
  units += (uint32_t)(NC * (ME_RA - R1));
  // B. the chunk's smallest partial sum; nothing at or below the bound: the chunk is done.  If promising, that candidate in full
  int total = ME_DENSE_T + 1;
  if (!dense_mode) {                   // (an earlier chunk of this macroblock had too many survivors: content without a sharp match)
    uint32_t m = a[0] * 16u;
#pragma unroll
    for (int j = 1; j < NC; j++) m = min(m, a[j] * 16u + (uint32_t)j);
    const uint32_t pm = __reduce_min_sync(0xffffffffu, xok ? (m >> 4) : 0xffffffffu);
    if (pm > bound) return nlist;
    if (2u * pm < bound) {                                             // (warp-uniform)
      const int src = __ffs(__ballot_sync(0xffffffffu, xok && (m >> 4) == pm)) - 1;
      const int cx = __shfl_sync(0xffffffffu, xi, src), cy = __shfl_sync(0xffffffffu, yb + (int)(m & 15u), src);
      bound = min(bound, coop_sad(win, s_cur, cx, cy, lane, lim));
      units += 1;                                                     // (2 packed SADs per lane, counted as a whole candidate row)
    }
    // C. survivors, counted per lane first (the ballots below are only for chunks that are compacted)
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < NC; j++) cnt += (a[j] <= bound) ? 1 : 0;
    total = __reduce_add_sync(0xffffffffu, xok ? cnt : 0);
    dense_mode = total > ME_DENSE_T;
  } else if (sweep_hopeless<NC>(a, xok, bound)) return nlist;
  if (total > ME_DENSE_T) {
    sweep_rows<NC, ME_RA, 16>(base, c, a, lim);
    units += (uint32_t)(NC * (16 - ME_RA));
    uint32_t r = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < NC; j++) r = min(r, a[j] * m2048 + (uint32_t)j);
    r = xok ? r + (uint32_t)(1 + xi * 32 + yb) : 0xffffffffu;
    best = min(best, r);
    bound = min(bound, __reduce_min_sync(0xffffffffu, r) >> 11);
    return nlist;
  }
  const uint32_t lt = (1u << lane) - 1u, tag = ((uint32_t)xi << 21) | ((uint32_t)yb << 16);
#pragma unroll
  for (int j = 0; j < NC; j++) {
    const bool sv = xok && a[j] <= bound;
    const uint32_t mk = __ballot_sync(0xffffffffu, sv);
    if (sv) SMW(list + nlist + __popc(mk & lt)) = a[j] + tag + ((uint32_t)j << 16);
    nlist += __popc(mk);
  }
  return nlist;
}

// tm_ref / tm_cur: u8 tensors {W, H, n_pairs}; boxes {48,47,1} and {16,16,1}.  grid = worker CTAs (persistent).
// VARIANT: ME_V_FULL exhaustive argmin straight from registers; ME_V_TSS the stock three-step search evaluating only
// the <= 33 positions it probes (8 candidates x 4 row quarters per step across the lanes, unaligned window words by
// funnel shift, no shifted copies); ME_V_SURF the whole surface in shared memory + a three-step walk over it (test
// hook, surface != nullptr: the two TSS implementations check each other).
template <int VARIANT>
__global__ void __launch_bounds__(ME_THREADS, 4)
me_search_kernel(const __grid_constant__ CUtensorMap tm_ref, const __grid_constant__ CUtensorMap tm_cur,
                 const __grid_constant__ MeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int WARP_WORDS = VARIANT == ME_V_SURF ? ME_WARP_WORDS_SURF : ME_WARP_WORDS_FULL;
  uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform for the compiler
  uint32_t* wsm = sm + warp * WARP_WORDS;                                // this warp's slice (128-byte aligned)
  uint32_t* s_pen = sm + ME_WARPS * WARP_WORDS;                          // [3][32] per row class: 0 or ME_ILLEGAL
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_pen + 96) + 2 * warp;   // one mbarrier per buffer
  uint32_t* s_sad = wsm + ME_SAD_OFF;
  // (debug build) what this thread may touch: its warp's slice, and the CTA's tables + mbarriers
  const SmLim lim = {smem_u32(wsm), smem_u32(wsm + WARP_WORDS), smem_u32(s_pen), smem_u32(s_pen + ME_CTA_WORDS)};
  (void)lim;

  if (threadIdx.x < 96) {
    const int cls = threadIdx.x >> 5, r = threadIdx.x & 31;
    const int lo = (a.rg.ylo >> (8 * cls)) & 0xff, hi = (int)((a.rg.yhi >> (8 * cls)) & 0xff) - 1;
    SMW(s_pen + threadIdx.x) = (r >= lo && r <= hi) ? 0u : ME_ILLEGAL;
  }
  if (lane == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }
  __syncthreads();

  const bool full = a.me_mode == P64B_ME_FULL;
  const int per_pair = a.mbw * a.mbh, total = per_pair * a.n_pairs;
  const int n_workers = gridDim.x * ME_WARPS;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.queue[a.parity ^ 1] = 0;      // the next launch's counter
  // macroblock n -> (pair, by, bx); n is also the raster index of its output record
  auto issue_loads = [&](int n, int buf) {
    const int pair = (int)(__umulhi((uint32_t)n, a.magic_pp) >> a.shift_pp), r = n - pair * per_pair;
    const int by = (int)(((uint32_t)r * a.magic_w) >> 16), bx = r - by * a.mbw;
    uint32_t* nb = wsm + buf * ME_BUF_WORDS;
    mbar_expect_tx(bars + buf, ME_WIN_BYTES + 256);
    tma_load_3d(nb, &tm_ref, bx * 16 - 16, by * 16 - 15, pair, bars + buf);      // out-of-frame samples arrive as zeros
    tma_load_3d(nb + 608, &tm_cur, bx * 16, by * 16, pair, bars + buf);
  };
  int n = blockIdx.x * ME_WARPS + warp;           // first macroblock: static; the rest come from the queue
  uint32_t ticket = 0;                            // lane 0: queue position of the warp's NEXT macroblock
  if (lane == 0 && n < total) issue_loads(n, 0);
  ticket = queue_take(a.queue + a.parity, lane);

  uint32_t units = 0;                             // candidate rows (16 pixels each) this warp's lanes actually accumulated
  for (int it = 0; n < total; it++) {
    // ---- the warp's next macroblock: start its loads into the other buffer, request the one after it
    const int b = it & 1;
    const int n_next = n_workers + (int)__shfl_sync(0xffffffffu, ticket, 0);
    if (lane == 0 && n_next < total) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's earlier generic-proxy reads are done
      issue_loads(n_next, b ^ 1);
    }
    if (n_next < total) ticket = queue_take(a.queue + a.parity, lane);       // (warp-uniform condition; lane 0 takes)
    // hint for the chunk order of the exhaustive search: the vertical vector this macroblock had in the output array before
    // this launch (in the encoder: the previous frame's vector of the same macroblock; anything else is harmless, the
    // order of evaluation never changes the result).  Loaded early, used after the window has arrived.
    int hint_my = 0;
    if (VARIANT == ME_V_FULL && lane == 0) hint_my = reinterpret_cast<const volatile int*>(a.out + n)[1];
    const int r_ = n - (int)(__umulhi((uint32_t)n, a.magic_pp) >> a.shift_pp) * per_pair;
    const int by = (int)(((uint32_t)r_ * a.magic_w) >> 16), bx = r_ - by * a.mbw;

    // ---- search geometry of this macroblock (warp-uniform, from the host table)
    const int cx = 8 * (bx == 0 ? 0 : (bx == a.mbw - 1 ? 2 : 1)), cyi = by == 0 ? 0 : (by == a.mbh - 1 ? 2 : 1);
    const int lxlo = (a.rg.xlo >> cx) & 0xff, lxhi = (int)((a.rg.xhi >> cx) & 0xff) - 1;
    const int lylo = (a.rg.ylo >> (8 * cyi)) & 0xff, lyhi = (int)((a.rg.yhi >> (8 * cyi)) & 0xff) - 1;
    const bool xr = lxhi - lxlo < 16;
    const int ndy = max(lyhi - lylo + 1, 0);
    const int rpg = xr ? (ndy + 1) >> 1 : ndy;                        // dy rows per lane group
    int xi = lxlo + (xr ? (lane & 15) : lane);
    const bool xok = xi <= lxhi;
    xi = min(xi, 30);                                                 // surplus lanes: any in-window column
    const int ystart = lylo + ((xr && lane >= 16) ? rpg : 0);
    const uint32_t* pen = s_pen + 32 * cyi;
    uint32_t* win = wsm + b * ME_BUF_WORDS;                           // copy 0 (TMA), [47][12]
    const uint32_t* s_cur = win + 608;                                // [16][4]
    uint32_t* shifted = wsm + ME_SHIFT_OFF + 8;                       // copy k at shifted + 584 (k-1)
    if (VARIANT == ME_V_SURF)
      for (int i = lane; i < 31 * 31; i += 32) SMW(s_sad + i) = 0xffffffffu;

    mbar_wait(bars + b, (it >> 1) & 1);
    // byte-shifted copies 1..3: each 16-byte quad + the following word yields the three shifted quads
#pragma unroll 1
    for (int u = lane; VARIANT != ME_V_TSS && u < ME_WIN_ROWS * 3; u += 32) {
      const int idx = 4 * u;                          // row (u/3) * 12 + 4 * (u%3)
      const uint4 v = SMR4(win + idx);
      const uint32_t nx = SMR(win + idx + 4);
      // funnel shift right by 8k as hi32(lo * 2^(32-8k)) + hi * 2^(32-8k): IMAD.HI + IMAD on the otherwise idle FMA pipe
      // (SHF would compete with VABSDIFF4 for the ALU pipe); the multipliers come in registers so they stay multiplies
#pragma unroll
      for (int k = 1; k < 4; k++) {
        const uint32_t m = k == 1 ? a.m24 : (k == 2 ? a.m16 : a.m8);
        uint4 o;
        o.x = __umulhi(v.x, m) + v.y * m; o.y = __umulhi(v.y, m) + v.z * m;
        o.z = __umulhi(v.z, m) + v.w * m; o.w = __umulhi(v.w, m) + nx * m;
        SMW4(shifted + (k - 1) * ME_COPY_WORDS + idx) = o;
      }
    }
    uint32_t c[16][4];
    if (VARIANT != ME_V_TSS) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint4 v = SMR4(s_cur + 4 * i);
        c[i][0] = v.x; c[i][1] = v.y; c[i][2] = v.z; c[i][3] = v.w;
      }
    }
    // SAD(0,0) = OMV (me.c:203, 271): two packed SADs per lane
    uint32_t omv;
    {
      const int i = lane >> 1, wc = (lane & 1) * 2;
      const uint2 r = SMR2(win + (15 + i) * ME_ROW_WORDS + 4 + wc);
      const uint2 cw = SMR2(s_cur + i * 4 + wc);
      omv = __reduce_add_sync(0xffffffffu, sad4(r.y, cw.y, sad4(r.x, cw.x, 0u)));
    }
    __syncwarp();

    // ---- SAD sweep over the legal rectangle
    uint32_t best = 0xffffffffu;
    if (VARIANT != ME_V_TSS) {
      const int o = xi + 1, k = o & 3;                             // o = dx + 16
      const uint32_t* colbase = (k ? shifted + (k - 1) * ME_COPY_WORDS : win) + (o >> 2);
      // exhaustive search: chunks of 10 dy rows (full_chunk).  Surface variant: passes of 10 (a last one of 5) dy rows.
      constexpr int PS = 10;
      const int np = rpg <= 0 ? 0 : (rpg - 1) / PS + 1;                 // passes k = 0..np-1 start at row PS*k
      uint32_t bound = omv;                                             // smallest full SAD so far
      if (VARIANT == ME_V_FULL) {
        // chunks of 10 dy rows, the one that holds the hinted dy (the macroblock's previous vector; 0 at first) first, then
        // outwards, so that the bound is good as early as possible (the order of evaluation never changes the result).  The
        // two-half mapping of edge columns keeps the plain order (the halves would disagree about the centre).  A last chunk of
        // fewer than 10 rows starts earlier instead (re-evaluating a few candidates: harmless for a minimum).
        uint32_t* list = s_sad;                                         // survivor list (the surface region of ME_V_SURF)
        int nlist = 0;
        bool dense_mode = false;
        const int hy = min(max(__shfl_sync(0xffffffffu, hint_my, 0), -15), 15) + 15;      // hinted dy as a surface row
        const int kc = !xr ? min(max((hy - lylo) / PS, 0), max(np - 1, 0)) : 0;
#pragma unroll 1
        for (int v = 0; v < 2 * np; v++) {
          const int k = !xr ? kc + ((v + 1) >> 1) * ((v & 1) ? -1 : 1) : v;
          if (k < 0 || k >= np) continue;
          if (nlist > ME_LIST_WORDS - ME_DENSE_T) {                     // a further chunk might not fit
            __syncwarp();
            sparse_finish(win, shifted, c, list, nlist, lane, a.m2048, bound, best, units, lim);
            __syncwarp();
            nlist = 0;
          }
          const int yb = min(ystart + max(min(PS * k, rpg - PS), 0), 22);
          nlist = full_chunk(colbase, win, s_cur, c, pen, list, nlist, xi, yb, xok, lane, a.m2048, bound, best, units, dense_mode, lim);
        }
        __syncwarp();
        sparse_finish(win, shifted, c, list, nlist, lane, a.m2048, bound, best, units, lim);
      } else {
        for (int v = 0; v < np; v++) {
          const int done = PS * v;
          uint32_t r;
          if (rpg - done > 5) r = sweep_pass<VARIANT, 10>(colbase, c, pen, s_sad, xi, min(ystart + done, 22), xok, a.m2048, units, lim);
          else                r = sweep_pass<VARIANT, 5>(colbase, c, pen, s_sad, xi, min(ystart + done, 27), xok, a.m2048, units, lim);
          best = min(best, (xok && r != 0xffffffffu) ? r + (uint32_t)(1 + xi * 32) : 0xffffffffu);
        }
      }
    }

    const size_t mb_index = (size_t)n;
    if (VARIANT == ME_V_SURF) {
      __syncwarp();
      if (lane == 0) SMW(s_sad + 15 * 31 + 15) = omv;     // (0,0) is in the surface even where a later probe of it would be illegal
      __syncwarp();
      if (a.surface)   // test hook: the whole surface, [dy+15][dx+15], 0xffffffff = illegal position
        for (int i = lane; i < 31 * 31; i += 32) a.surface[mb_index * 961 + i] = SMR(s_sad + i);
    }

    // ---- search result
    int mx = 0, my = 0;
    uint32_t mv = omv;
    if (full) {
      // FastBME (me.c:206-227) scans dx outer, dy inner with strict <, after probing (0,0): winner = min of
      // SAD<<11 | (1 + (dx+15)*32 + (dy+15)); (0,0) enters with order 0 so that it wins ties
      best = min(__reduce_min_sync(0xffffffffu, best), omv << 11);       // `best` already carries the whole key
      mv = best >> 11;
      const int ord = best & 2047;
      if (ord) { mx = ((ord - 1) >> 5) - 15; my = ((ord - 1) & 31) - 15; }
    } else if (VARIANT == ME_V_TSS) {
      // StepBME (me.c:273-311) evaluating only what it probes: steps 8,4,2,1; the 8 neighbours of the centre in
      // (diry outer, dirx inner) order = candidate slot lane>>2; lane&3 = which four rows of the block this lane sums.
      // Legality me.c:292-294; strict < keeps the earlier candidate on ties and the centre on equality.
      const int cand = lane >> 2, qr = lane & 3, nn = cand < 4 ? cand : cand + 1;
      uint32_t cq[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint4 v = SMR4(s_cur + 4 * (4 * qr + i));
        cq[i][0] = v.x; cq[i][1] = v.y; cq[i][2] = v.z; cq[i][3] = v.w;
      }
      for (int step = 8; step >= 1; step >>= 1) {
        const int dxi = mx + (nn % 3 - 1) * step + 15, dyi = my + (nn / 3 - 1) * step + 15;
        const bool legal = dxi >= lxlo && dxi <= lxhi && dyi >= lylo && dyi <= lyhi;
        uint32_t sad = 0;
        if (legal) {
          const int o = dxi + 1, sh = (o & 3) * 8;
          const uint32_t* p = win + (dyi + 4 * qr) * ME_ROW_WORDS + (o >> 2);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t w0 = SMR(p + i * ME_ROW_WORDS), w1 = SMR(p + i * ME_ROW_WORDS + 1), w2 = SMR(p + i * ME_ROW_WORDS + 2),
                           w3 = SMR(p + i * ME_ROW_WORDS + 3), w4 = SMR(p + i * ME_ROW_WORDS + 4);
            sad = sad4(__funnelshift_r(w0, w1, sh), cq[i][0], sad);
            sad = sad4(__funnelshift_r(w1, w2, sh), cq[i][1], sad);
            sad = sad4(__funnelshift_r(w2, w3, sh), cq[i][2], sad);
            sad = sad4(__funnelshift_r(w3, w4, sh), cq[i][3], sad);
          }
        }
        sad += __shfl_xor_sync(0xffffffffu, sad, 1);
        sad += __shfl_xor_sync(0xffffffffu, sad, 2);
        const uint32_t key = legal ? min((sad << 4) | (uint32_t)(cand + 1), mv << 4) : (mv << 4);
        const uint32_t bk = __reduce_min_sync(0xffffffffu, key);
        const int ord = bk & 15;
        if (ord) {
          const int n2 = ord - 1 < 4 ? ord - 1 : ord;
          mx += (n2 % 3 - 1) * step; my += (n2 / 3 - 1) * step;
          mv = bk >> 4;
        }
      }
    } else if (VARIANT == ME_V_SURF) {
      // StepBME (me.c:273-311): steps 8,4,2,1; 8 neighbours in (diry outer, dirx inner) order; the centre
      // moves once per step; strict < keeps the earlier candidate on ties.  Lanes 0..7 probe in parallel.
      for (int step = 8; step >= 1; step >>= 1) {
        uint32_t key = (mv << 4);                     // current best, order 0
        if (lane < 8) {
          int nn = lane < 4 ? lane : lane + 1;        // skip the centre (index 4)
          int dx = mx + (nn % 3 - 1) * step, dy = my + (nn / 3 - 1) * step;
          if (dx >= -15 && dx <= 15 && dy >= -15 && dy <= 15) {
            uint32_t sv = SMR(s_sad + (dy + 15) * 31 + dx + 15);
            if (sv != 0xffffffffu) key = min(key, (sv << 4) | (uint32_t)(lane + 1));
          }
        }
        const uint32_t bk = __reduce_min_sync(0xffffffffu, key);
        int ord = bk & 15;
        if (ord) {
          int nn = ord - 1; nn = nn < 4 ? nn : nn + 1;
          mx += (nn % 3 - 1) * step; my += (nn / 3 - 1) * step;
          mv = bk >> 4;
        }
      }
    }

    // ---- statistics over the best-match reference block (me.c:230-245), on packed words, two per lane:
    // sum r = SAD(r,0); sum r^2 = dp4a(r,r); sum (r-c)^2 = dp4a(r,r) - 2 dp4a(r,c) + dp4a(c,c)
    {
      const int i = lane >> 1, wc = (lane & 1) * 2, oo = mx + 16, sh = (oo & 3) * 8;
      const uint32_t* rp = win + (my + 15 + i) * ME_ROW_WORDS + (oo >> 2) + wc;      // unaligned: three words, two funnel shifts
      const uint32_t w0 = SMR(rp), w1 = SMR(rp + 1), w2 = SMR(rp + 2);
      const uint32_t ra = __funnelshift_r(w0, w1, sh), rb = __funnelshift_r(w1, w2, sh);
      const uint2 cw = SMR2(s_cur + i * 4 + wc);
      uint32_t smm = sad4(rb, 0u, sad4(ra, 0u, 0u));
      uint32_t so = __dp4a(rb, rb, __dp4a(ra, ra, 0u));
      uint32_t sv = so + __dp4a(cw.y, cw.y, __dp4a(cw.x, cw.x, 0u)) - 2u * __dp4a(rb, cw.y, __dp4a(ra, cw.x, 0u));
      sv = __reduce_add_sync(0xffffffffu, sv);
      so = __reduce_add_sync(0xffffffffu, so);
      smm = __reduce_add_sync(0xffffffffu, smm);
      if (lane == 0) {
        const int var = (int)sv / 256, mwor = (int)smm;
        const int varor = (int)so / 256 - (mwor / 256) * (mwor / 256);
        int4* op = reinterpret_cast<int4*>(a.out + mb_index);
        op[0] = make_int4(mx, my, (int)mv, (int)omv);
        op[1] = make_int4(var, varor, mwor, 0);
      }
    }
    __syncwarp();          // every lane is done with this buffer and the shifted copies
    n = n_next;
  }
  // executed work of this launch, in warp-wide candidate rows (x 32 lanes x 4 packed SADs): the roofline's numerator
  if (a.executed && lane == 0) atomicAdd(a.executed, (unsigned long long)units);
}

// ---------------------------------------------------------------------------------------------------
// Transform chain, one 8x8 block per thread, everything in registers
// ---------------------------------------------------------------------------------------------------
#define P64B_MS(e) ((e) >> 9)     // MSCALE: arithmetic shift = floor (chendct.c:46-53, NO_MULTIPLY)

// Chen forward butterfly on 8 values (chendct.c:119-158 / 165-197). PRE: 1 = column pass (<<2), 0 = row pass (>>1),
// 2 = column pass on inputs that already carry the <<2 (exact: the shift distributes over the sums)
template <int PRE>
__device__ __forceinline__ void fdct8(int& x0, int& x1, int& x2, int& x3, int& x4, int& x5, int& x6, int& x7) {
  int a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3;
  if (PRE == 2) {
    a0 = x0 + x7; c3 = x0 - x7; a1 = x1 + x6; c2 = x1 - x6;
    a2 = x2 + x5; c1 = x2 - x5; a3 = x3 + x4; c0 = x3 - x4;
  } else if (PRE) {
    a0 = (x0 + x7) << 2; c3 = (x0 - x7) << 2; a1 = (x1 + x6) << 2; c2 = (x1 - x6) << 2;
    a2 = (x2 + x5) << 2; c1 = (x2 - x5) << 2; a3 = (x3 + x4) << 2; c0 = (x3 - x4) << 2;
  } else {
    a0 = (x0 + x7) >> 1; c3 = (x0 - x7) >> 1; a1 = (x1 + x6) >> 1; c2 = (x1 - x6) >> 1;
    a2 = (x2 + x5) >> 1; c1 = (x2 - x5) >> 1; a3 = (x3 + x4) >> 1; c0 = (x3 - x4) >> 1;
  }
  b0 = a0 + a3; b1 = a1 + a2; b2 = a1 - a2; b3 = a0 - a3;
  x0 = P64B_MS(362 * (b0 + b1));
  x4 = P64B_MS(362 * (b0 - b1));
  x2 = P64B_MS(196 * b2 + 473 * b3);
  x6 = P64B_MS(196 * b3 - 473 * b2);
  b0 = P64B_MS(362 * (c2 - c1));
  b1 = P64B_MS(362 * (c2 + c1));
  a0 = c0 + b0; a1 = c0 - b0; a2 = c3 - b1; a3 = c3 + b1;
  x1 = P64B_MS(100 * a0 + 502 * a3);
  x3 = P64B_MS(426 * a2 - 284 * a1);
  x5 = P64B_MS(426 * a1 + 284 * a2);
  x7 = P64B_MS(100 * a3 - 502 * a0);
}

// Chen inverse butterfly (chendct.c:236-299 / 305-363). PRE: 1 = column pass (<<2); 2 = column pass on inputs that
// already carry the <<2
template <int PRE>
__device__ __forceinline__ void idct8(int& x0, int& x1, int& x2, int& x3, int& x4, int& x5, int& x6, int& x7) {
  const int sh = PRE == 1 ? 2 : 0;
  int b0 = x0 << sh, a0 = x1 << sh, b2 = x2 << sh, a1 = x3 << sh;
  int b1 = x4 << sh, a2 = x5 << sh, b3 = x6 << sh, a3 = x7 << sh;
  int c0 = P64B_MS(100 * a0 - 502 * a3);
  int c1 = P64B_MS(426 * a2 - 284 * a1);
  int c2 = P64B_MS(426 * a1 + 284 * a2);
  int c3 = P64B_MS(502 * a0 + 100 * a3);
  a0 = P64B_MS(362 * (b0 + b1));
  a1 = P64B_MS(362 * (b0 - b1));
  a2 = P64B_MS(196 * b2 - 473 * b3);
  a3 = P64B_MS(473 * b2 + 196 * b3);
  b0 = a0 + a3; b1 = a1 + a2; b2 = a1 - a2; b3 = a0 - a3;
  a0 = c0 + c1; a1 = c0 - c1; a2 = c3 - c2; a3 = c3 + c2;
  c0 = a0; c1 = P64B_MS(362 * (a2 - a1)); c2 = P64B_MS(362 * (a2 + a1)); c3 = a3;
  x0 = b0 + c3; x1 = b1 + c2; x2 = b2 + c1; x3 = b3 + c0;
  x4 = b3 - c0; x5 = b2 - c1; x6 = b1 - c2; x7 = b0 - c3;
}

// trunc((v<0 ? v-h : v+h) / 2h) for h = 4, 8 (chendct.c:205, 374) == (v + h + (v>>31)) >> log2(2h)   [arithmetic shifts]
__device__ __forceinline__ int round_div8(int v) { return (v + 4 + (v >> 31)) >> 3; }
__device__ __forceinline__ int round_div16(int v) { return (v + 8 + (v >> 31)) >> 4; }

// raster index of the coefficient that lands at zig-zag position k (inverse of transform.c:67-75)
__host__ __device__ constexpr int izig(int k) {
  constexpr int t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return t[k];
}
// transmission position of raster coefficient i: ZigzagMatrix is the scatter out[zz[i]] = in[i] (transform.c:561-568)
struct ZigTable { uint8_t pos[64]; };
__host__ __device__ constexpr ZigTable make_zig() {
  ZigTable z{};
  for (int k = 0; k < 64; k++) z.pos[izig(k)] = (uint8_t)k;
  return z;
}
__constant__ ZigTable c_zig = make_zig();
__host__ __device__ constexpr int c_zig_at(int i) { return make_zig().pos[i]; }

__device__ __forceinline__ int ubyte(uint32_t w, int k) { return (int)__byte_perm(w, 0, 0x4440 + k); }   // PRMT, zero-extended byte k
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {                                   // low bytes of a,b,c,d
  return __byte_perm(__byte_perm((uint32_t)a, (uint32_t)b, 0x0040), __byte_perm((uint32_t)c, (uint32_t)d, 0x0040), 0x5410);
}

// four ints -> four bytes (a lowest) with saturation to [0,255]: two I2IP (cvt.pack.sat) instead of 4 x min/max + 3 PRMT
__device__ __forceinline__ uint32_t pack4_sat(int a, int b, int c, int d) {
  uint32_t hi, r;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;
}

// MType property tables (p64.c:217-222) as bit masks over type 0..9
constexpr uint32_t M_CBP = 0x36c, M_INTRA = 0x003, M_MF = 0x3f0, M_FILTER = 0x380, M_TCOEF = 0x36f;
__device__ __forceinline__ bool mt_is(uint32_t mask, int mt) { return (mask >> mt) & 1u; }

struct MbArgs {
  Geom g;
  const uint8_t* src;      // [S][frame_bytes] source frames
  const uint8_t* ref;      // [S][frame_bytes] previous reconstruction (CFS)
  uint8_t* out;            // [S][frame_bytes] reconstruction being written (OFS)
  const p64b_me* me;       // [S][nmb raster]
  const uint8_t* li_prev;  // [S][nmb] LastIntra before this frame
  uint8_t* li_new;         // [S][nmb]
  const uint8_t* quant;    // [S] per-stream GQUANT or nullptr
  p64b_mb* mbs;            // output records
  int8_t* levels;          // output levels
  int n_streams;
  int gob_first, gob_count;   // task t -> stream t / gob_count, GOB gob_first + t % gob_count
  int out_mb_per_stream;      // stride of the output arrays per stream, in MBs (nmb or 33)
  int first_frame, force_intra, gquant;
  uint32_t mps_magic, mps_shift;  // n / (gob_count*33) = umulhi(n, mps_magic) >> mps_shift
};

// One 8-byte row of the prediction at any byte alignment: aligned 8-byte loads + funnel shift
// (frame stores carry slack at the end).
__device__ __forceinline__ uint2 fetch_row8(const uint8_t* p) {
  const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
  const uint2* qp = reinterpret_cast<const uint2*>(ad & ~(uintptr_t)7);
  const int sh = (int)(ad & 7) * 8;
  const uint2 r0 = __ldg(qp);
  if (sh == 0) return r0;
  const uint2 r1 = __ldg(qp + 1);
  uint32_t w0 = r0.x, w1 = r0.y, w2 = r1.x;
  if (sh & 32) { w0 = r0.y; w1 = r1.x; w2 = r1.y; }
  return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// ---------------------------------------------------------------------------------------------------
// Transform chain: one thread per 8x8 block with the block in a private shared-memory tile.
// CTA = 6 warps x 32 macroblocks (warp c = block index c of p64.c:77-79, lane = macroblock): a warp is homogeneous
// in plane handling.  The column passes (first in the reference, chendct.c:119-158 / 236-299) run on half blocks --
// 8 rows x 4 columns = 32 registers -- read and written as 16-byte rows of the tile; the row passes stream one row at
// a time.  That keeps the thread at <= 80 registers (24 warps per SM) without the transposes and per-lane overheads of
// the 8-lanes-per-block layout, and the instruction count at the one-thread-per-block level (~3 700 per block).
// Tile stride 84 words (== 20 mod 32): a quarter warp's 16-byte row accesses fall into 8 different 4-bank groups.
// ---------------------------------------------------------------------------------------------------
constexpr int MB4_PER_CTA = 32;
constexpr int MB4_THREADS = 6 * MB4_PER_CTA;
constexpr int MB4_TILE = 84;                      // words per thread: [8][8] int32 block + 16 words packed prediction + 4 pad
constexpr int MB4_SMEM = MB4_THREADS * MB4_TILE * 4;

// sum_j ubyte(a, j) * sbyte(b, j) + c  (IDP.4A.U8.S8): with b = k << 8j this is k * (byte j of a) + c on the FMA pipe
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// byte 0 of v into byte position b of w (PRMT, selector known at compile time)
template <int B>
__device__ __forceinline__ uint32_t put_byte(uint32_t w, int v) {
  return __byte_perm(w, (uint32_t)v, B == 0 ? 0x3214 : (B == 1 ? 0x3240 : (B == 2 ? 0x3410 : 0x4210)));
}
template <int K>
__device__ __forceinline__ void zig_insert(uint32_t (&out)[16], int v) { out[K >> 2] = put_byte<K & 3>(out[K >> 2], v); }

// (debug build) the calling thread's own tile in the dynamic shared memory of mb_encode_kernel / mb_decode_kernel
__device__ __forceinline__ SmLim mb_tile_lim() {
#ifdef P64B_BOUNDS_CHECK
  extern __shared__ __align__(16) uint32_t dbg_dyn[];
  const uint32_t lo = smem_u32(dbg_dyn) + threadIdx.x * MB4_TILE * 4u;
  return SmLim{lo, lo + MB4_TILE * 4u, lo, lo + MB4_TILE * 4u};
#else
  return SmLim{0, 0, 0, 0};
#endif
}
__device__ __forceinline__ void st_row(int* tile, int r, int h, const int (&v)[4]) {
  const SmLim lim = mb_tile_lim(); (void)lim;
  P64B_SM_T(int4, tile + 8 * r + 4 * h, lim) = make_int4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void ld_row(const int* tile, int r, int h, int (&v)[4]) {
  const SmLim lim = mb_tile_lim(); (void)lim;
  const int4 q = P64B_SM_T(const int4, tile + 8 * r + 4 * h, lim);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

// forward row pass of row R: Chen row butterfly, then ChenDct's final rounding + BoundDctMatrix + CCITT[Flat]Quantize +
// [Flat]BoundQuantizeMatrix (chendct.c:204-205, transform.c:271-350, 460-537) as ONE signed multiply-shift on the RAW output v:
//   |x| = (|v|+4)>>3 clamped to 1023  ==  (min(|v|,8187)+4)>>3
//   |level| = min(127, floor((|x| + ev) / 2Q)) = min(127, (a*M + K) >> 22),  a = min(|v|,8187), M = floor(2^22/16Q)+1,
//             K = (4+8ev)*M                                                  (ev = 1 for even Q; nested floors)
//   both clamps become one clamp of v to +-A(Q), A = the largest a <= 8187 whose level is <= 127, and the sign moves into
//   the rounding constant:  level = (clamp(v,-A,A)*M + (v < 0 ? K ^ 0x3fffff : K)) >> 22   (arithmetic shift)
// because -((a*M+K) >> 22) == (-a*M + (2^22-1-K)) >> 22 and K < 2^22 (tests/test_abi_and_host.py checks the identity
// for every raw value and every Q).  Sum |level| is only ever compared with 0 and 1 (p64.c:887-903); sum level^2 gives the
// same two answers and is one IMAD per coefficient.  Levels go back into the tile, their low bytes into the
// transmission-order words.
template <int R>
__device__ __forceinline__ int fwd_row(int* tile, uint32_t (&out)[16], int M, int K, int A, uint32_t ev, bool intra) {
  int v[8];
  {
    int a[4], b[4];
    ld_row(tile, R, 0, a); ld_row(tile, R, 1, b);
    v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; v[3] = a[3]; v[4] = b[0]; v[5] = b[1]; v[6] = b[2]; v[7] = b[3];
  }
  fdct8<0>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
  int l[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (R == 0 && j == 0) {                                   // the DC term (clamped from above only, transform.c:464-466)
      const int x = min(round_div8(v[0]), 2047);
      int d;
      if (intra) d = min(max((x + 4) >> 3, 1), 254);          // (x-4)/8 <= 0 for x <= 0 and is clamped to 1 anyway
      else { d = min((int)(((uint32_t)(abs(x) + (int)ev) * (uint32_t)M) >> 19), 127); d = x < 0 ? -d : d; }
      l[0] = d;
    } else {
      const int x = v[j], sg = x >> 31;
      l[j] = (min(max(x, -A), A) * M + (K ^ (sg & 0x3fffff))) >> 22;
    }
    sum += l[j] * l[j];
  }
  zig_insert<c_zig_at(8 * R + 0)>(out, l[0]); zig_insert<c_zig_at(8 * R + 1)>(out, l[1]);
  zig_insert<c_zig_at(8 * R + 2)>(out, l[2]); zig_insert<c_zig_at(8 * R + 3)>(out, l[3]);
  zig_insert<c_zig_at(8 * R + 4)>(out, l[4]); zig_insert<c_zig_at(8 * R + 5)>(out, l[5]);
  zig_insert<c_zig_at(8 * R + 6)>(out, l[6]); zig_insert<c_zig_at(8 * R + 7)>(out, l[7]);
  {
    const int a[4] = {l[0], l[1], l[2], l[3]}, b[4] = {l[4], l[5], l[6], l[7]};
    st_row(tile, R, 0, a); st_row(tile, R, 1, b);
  }
  return sum;
}

// H.261 loop filter (LoadFilterMatrix, io.c:323-372) of one output row on packed bytes: the separable 1-2-1 filter with
// block-edge rows / columns passed through is one 2-D kernel with a single rounding (S16+8)>>4 (identity checked in
// tests/test_oracle_vs_ref.py), so every output sample is a chain of byte dot products (IDP.4A, FMA pipe) over the row
// above, the row and the row below with the column weights folded into the coefficient bytes.  wa/wb/wc = vertical
// weights of the three rows ((1,2,1) inside the block, (0,4,0) on its first and last row).
template <int WA, int WB, int WC>
__device__ __forceinline__ void filter_row(const uint32_t (&pk)[16], int ra, int rb, int rc, uint32_t& o0, uint32_t& o1) {
  // horizontal coefficient bytes per output column j for the two words of a row (columns 0-3, 4-7)
  constexpr int C0[8] = {0x00000004, 0x00010201, 0x01020100, 0x02010000, 0x01000000, 0, 0, 0};
  constexpr int C1[8] = {0, 0, 0, 0x00000001, 0x00000102, 0x00010201, 0x01020100, 0x04000000};
  int o[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int acc = 8;
    if (WA) { if (C0[j]) acc = dp4a_us(pk[2 * ra], C0[j] * WA, acc); if (C1[j]) acc = dp4a_us(pk[2 * ra + 1], C1[j] * WA, acc); }
    if (WB) { if (C0[j]) acc = dp4a_us(pk[2 * rb], C0[j] * WB, acc); if (C1[j]) acc = dp4a_us(pk[2 * rb + 1], C1[j] * WB, acc); }
    if (WC) { if (C0[j]) acc = dp4a_us(pk[2 * rc], C0[j] * WC, acc); if (C1[j]) acc = dp4a_us(pk[2 * rc + 1], C1[j] * WC, acc); }
    o[j] = acc >> 4;
  }
  o0 = pack4_sat(o[0], o[1], o[2], o[3]); o1 = pack4_sat(o[4], o[5], o[6], o[7]);
}

// per quantiser: the multiplier M and the clamp A(Q) of fwd_row (compile-time table in constant memory: a stream's 32
// macroblocks of a warp share the quantiser, so the access is uniform)
struct QuantTable { uint32_t m[32], a[32]; };
__host__ __device__ constexpr QuantTable make_quant_table() {
  QuantTable t{};
  for (uint32_t qq = 1; qq < 32; qq++) {
    const uint32_t m = (1u << 18) / qq + 1u, k = ((qq & 1) ? 4u : 12u) * m, a = ((1u << 29) - 1u - k) / m;
    t.m[qq] = m; t.a[qq] = a < 8187u ? a : 8187u;
  }
  return t;
}
__constant__ QuantTable c_quant = make_quant_table();

__global__ void __launch_bounds__(MB4_THREADS, 3)
mb_encode_kernel(const __grid_constant__ MbArgs a) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ int s_acc[6][MB4_PER_CTA];
  const Geom& g = a.g;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;        // block index within the MB: p64.c:77-79
  int* tile = reinterpret_cast<int*>(s_dyn) + threadIdx.x * MB4_TILE;
  uint32_t* s_pk = reinterpret_cast<uint32_t*>(tile) + 64;        // packed prediction, 8 rows x 2 words
  SmLim lim = mb_tile_lim();                                      // (debug build) + the CTA's exchange array
  lim.lo2 = smem_u32(&s_acc[0][0]); lim.hi2 = lim.lo2 + (uint32_t)sizeof(s_acc);
  (void)lim;

  const int n_total = a.n_streams * a.gob_count * 33;
  const int n = blockIdx.x * MB4_PER_CTA + lane;
  const bool active = n < n_total;
  const int nn = active ? n : n_total - 1;
  const int mps = a.gob_count * 33;
  const int s = (int)(__umulhi((uint32_t)nn, a.mps_magic) >> a.mps_shift), rem = nn - s * mps;
  const int gob = a.gob_first + rem / 33, m = rem % 33;
  int col, row;                                                    // MoveTo, io.c:730-741
  if (g.qcif) { col = m % 11; row = gob * 3 + m / 11; }
  else { col = (gob & 1) * 11 + m % 11; row = (gob >> 1) * 3 + m / 11; }
  const int mbi = gob * 33 + m;                                    // GOB-major index
  const bool chroma = c >= 4;
  const int w = chroma ? g.W >> 1 : g.W, wq = w >> 3;
  const int off = chroma ? g.W * g.H + (c - 4) * (g.W * g.H >> 2) + row * 8 * w + col * 8
                         : (row * 16 + (c >> 1) * 8) * w + col * 16 + (c & 1) * 8;
  const size_t fo = (size_t)s * g.frame_bytes;

  // ---- source block (ReadBlock, io.c:793-820): issued first, it does not depend on the decision
  uint2 srow[8];
  {
    const uint2* sp = reinterpret_cast<const uint2*>(a.src + fo + off);
#pragma unroll
    for (int i = 0; i < 8; i++) srow[i] = __ldg(sp + i * wq);
  }

  // ---- MTYPE decision (p64.c:734-773), double arithmetic exactly as written
  int4 me0 = make_int4(0, 0, 0, 0), me1 = me0;
  if (!a.first_frame) {
    const int4* mp = reinterpret_cast<const int4*>(a.me + (size_t)s * g.mbw * g.mbh + row * g.mbw + col);
    me0 = __ldg(mp); me1 = __ldg(mp + 1);
  }
  const int mvx = me0.x, mvy = me0.y;
  const int li = a.li_prev[(size_t)s * g.nmb + mbi];
  // (every warp takes its macroblocks' decisions itself: 40 instructions, cheaper than a CTA-wide barrier to share them)
  int mt = 0;
  if (!a.first_frame && !a.force_intra) {
    const int oval = me0.w, val = me0.z, var = me1.x, varor = me1.y;
    if (var < 64 || varor > var) {
      // x = OVal/256.0, y = Val/256.0 (exact in double): x < 1.0 <=> OVal < 256; x < 3.0 <=> OVal < 768; y > x*0.5 <=> 2 Val > OVal
      // (exact products); only y > x/1.1 needs the IEEE division
      const double x = (double)oval / 256.0, y = (double)val / 256.0;
      if (oval < 256 || (oval < 768 && 2 * val > oval) || y > __ddiv_rn(x, 1.1)) mt = 2;
      else if (var < 6) mt = 5;
      else mt = 8;
    }
  }
  if (li > 131) mt = 0;
  const int q = a.quant ? a.quant[s] : a.gquant;
  const bool intra = mt_is(M_INTRA, mt);
  const uint32_t ev = (q & 1) ? 0u : 1u;
  const int M = (int)c_quant.m[q], K = (int)(4u + 8u * ev) * M, A = (int)c_quant.a[q];

  // ---- prediction (SubOverlay / SubCompensate / HalfSubCompensate addressing, io.c:142-313; chroma vector = MV/2 with C
  // truncation, io.c:268-269), as packed bytes: pk[2r], pk[2r+1] = row r
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; i++) pk[i] = 0;
  auto fetch_pred = [&](bool mc) {
    int dx = 0, dy = 0;
    if (mc) { dx = chroma ? mvx / 2 : mvx; dy = chroma ? mvy / 2 : mvy; }
    const uint8_t* b = a.ref + fo + off + dy * w + dx;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint2 r = fetch_row8(b + i * w); pk[2 * i] = r.x; pk[2 * i + 1] = r.y; }
  };
  if (!intra) {
    fetch_pred(mt_is(M_MF, mt));
    if (mt_is(M_FILTER, mt)) {
      uint32_t f[16];
      filter_row<0, 4, 0>(pk, 0, 0, 0, f[0], f[1]);
      filter_row<1, 2, 1>(pk, 0, 1, 2, f[2], f[3]);   filter_row<1, 2, 1>(pk, 1, 2, 3, f[4], f[5]);
      filter_row<1, 2, 1>(pk, 2, 3, 4, f[6], f[7]);   filter_row<1, 2, 1>(pk, 3, 4, 5, f[8], f[9]);
      filter_row<1, 2, 1>(pk, 4, 5, 6, f[10], f[11]); filter_row<1, 2, 1>(pk, 5, 6, 7, f[12], f[13]);
      filter_row<0, 4, 0>(pk, 7, 7, 7, f[14], f[15]);
#pragma unroll
      for (int i = 0; i < 16; i++) pk[i] = f[i];
    }
  }

  // ---- ReadCompressMDU (p64.c:823-886): residual + Chen column pass on half blocks, into the tile
#pragma unroll
  for (int h = 0; h < 2; h++) {
    int v[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      // 4 (s_j - p_j) by two byte dot products (IDP.4A, FMA pipe) instead of two byte extractions, a subtraction and the
      // column pass's <<2 on the ALU pipe, which is this kernel's bottleneck
      const uint32_t sw = h ? srow[r].y : srow[r].x, pw = pk[2 * r + h];
#pragma unroll
      for (int j = 0; j < 4; j++) v[r][j] = dp4a_us(sw, 4 << (8 * j), dp4a_us(pw, (int)(0xfcu << (8 * j)), 0));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) fdct8<2>(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
    for (int r = 0; r < 8; r++) st_row(tile, r, h, v[r]);
  }
  // the prediction moves to shared memory: it is needed again only row by row in the reconstruction
#pragma unroll
  for (int i = 0; i < 4; i++) SMW4(s_pk + 4 * i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);

  // ---- row pass, bound, quantise, zig-zag (transform.c:561-568): byte k of the output = level at raster izig(k)
  uint32_t out[16];
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = 0;
  int acc = 0;
  acc += fwd_row<0>(tile, out, M, K, A, ev, intra); acc += fwd_row<1>(tile, out, M, K, A, ev, intra);
  acc += fwd_row<2>(tile, out, M, K, A, ev, intra); acc += fwd_row<3>(tile, out, M, K, A, ev, intra);
  acc += fwd_row<4>(tile, out, M, K, A, ev, intra); acc += fwd_row<5>(tile, out, M, K, A, ev, intra);
  acc += fwd_row<6>(tile, out, M, K, A, ev, intra); acc += fwd_row<7>(tile, out, M, K, A, ev, intra);
  const size_t mb_out = (size_t)s * a.out_mb_per_stream + (mbi - a.gob_first * 33);
  if (active) {
    uint4* lp = reinterpret_cast<uint4*>(a.levels + mb_out * 384 + c * 64);
#pragma unroll
    for (int i = 0; i < 4; i++) lp[i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
  }

  // ---- CBP and the type-4 / type-7 fallback (p64.c:887-908)
  P64B_SM_T(int, &s_acc[c][lane], lim) = acc;
  __syncthreads();
  int cbp = 0x3f, nz = 0;
  {
    int pm = 0, cb = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int ak = P64B_SM_T(const int, &s_acc[k][lane], lim);
      if (ak && !pm) pm = 1 << (5 - k);
      if (ak > 1) cb |= 1 << (5 - k);
      if (ak) nz |= 1 << (5 - k);
    }
    if (mt_is(M_CBP, mt)) {
      cbp = cb ? cb : pm;
      if (!cbp) { mt = mt_is(M_FILTER, mt) ? 7 : 4; cbp = 0x3f; }     // no coefficients at all: levels are already zero
    }
  }
  const int mt_final = mt;

  // ---- inverse half (p64.c:935-959) + DecodeSaveMDU (p64.c:971-1013)
  const bool coded = ((cbp >> (5 - c)) & 1) && mt_is(M_TCOEF, mt_final);
  // a type-2 MB that fell back to type 4 predicts with the ME vector (p64.c:904, marker.c:339-342)
  if (!intra && mt_final == 4 && (mvx | mvy)) {
    fetch_pred(true);
#pragma unroll
    for (int i = 0; i < 4; i++) SMW4(s_pk + 4 * i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
  uint2* op = reinterpret_cast<uint2*>(a.out + fo + off);
  if (coded) {
    // inverse quantise (transform.c:359-451): (2|l|+1)Q - ev with the sign of l, 0 stays 0; intra DC = 8 l; Chen column pass
    // (the column pass's <<2 is folded into the constants)
    const int q2 = 8 * q, qo = 4 * (q - (int)ev);
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      int v[8][4];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        ld_row(tile, r, h, v[r]);
#pragma unroll
        for (int j = 0; j < 4; j++) {              // 4 [(2|l|+1) Q - ev] sign(l) = l * 8Q + sign(l) * 4 (Q - ev); sign(0) = 0
          const int l = v[r][j];
          v[r][j] = l * q2 + min(max(l, -1), 1) * qo;
        }
      }
      if (h == 0 && intra) v[0][0] = ((const int*)tile)[0] * 32;
#pragma unroll
      for (int j = 0; j < 4; j++) idct8<2>(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
      __syncwarp();                                 // (no cross-thread hazard: keeps the loads of the rolled loop above the stores)
#pragma unroll
      for (int r = 0; r < 8; r++) st_row(tile, r, h, v[r]);
    }
    // row pass, ChenIDct rounding + Add*Compensate + BoundIDctMatrix, row by row
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
      int x[8];
      {
        int p0[4], p1[4];
        ld_row(tile, r, 0, p0); ld_row(tile, r, 1, p1);
        x[0] = p0[0]; x[1] = p0[1]; x[2] = p0[2]; x[3] = p0[3]; x[4] = p1[0]; x[5] = p1[1]; x[6] = p1[2]; x[7] = p1[3];
      }
      idct8<0>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
      const uint2 pw = SMR2(s_pk + 2 * r);
      int o[8];
#pragma unroll
      for (int j = 0; j < 4; j++) {              // + prediction byte j by a byte dot product (FMA pipe); the clamp is in the pack
        o[j] = dp4a_us(pw.x, 1 << (8 * j), round_div16(x[j]));
        o[4 + j] = dp4a_us(pw.y, 1 << (8 * j), round_div16(x[4 + j]));
      }
      if (active) op[r * wq] = make_uint2(pack4_sat(o[0], o[1], o[2], o[3]), pack4_sat(o[4], o[5], o[6], o[7]));
    }
  } else if (active) {
#pragma unroll
    for (int r = 0; r < 8; r++) op[r * wq] = SMR2(s_pk + 2 * r);   // reconstruction = prediction
  }
  if (active && c == 0) {
    const bool mf = mt_is(M_MF, mt_final);
    uint32_t r0 = (uint32_t)mt_final | ((uint32_t)cbp << 8) | ((uint32_t)((mf ? mvx : 0) & 0xff) << 16) |
                  ((uint32_t)((mf ? mvy : 0) & 0xff) << 24);
    uint32_t r1 = (uint32_t)q | ((uint32_t)nz << 8);
    *reinterpret_cast<uint2*>(a.mbs + mb_out) = make_uint2(r0, r1);
    a.li_new[(size_t)s * g.nmb + mbi] = mt_is(M_INTRA, mt_final) ? 0 : (uint8_t)(li + 1);   // p64.c:909-910
  }
}

// ---------------------------------------------------------------------------------------------------
// Decoder's inverse half (DecompressMDU p64.c:1179-1237 + DecodeSaveMDU p64.c:971-1013): the same
// inverse quantise / Chen IDCT / prediction (overlay, MC, half-vector chroma, loop filter) / clamp / store as the
// encoder's reconstruction, driven by macroblock records and levels a bit-stream PARSER produced (decoder.cpp) instead of
// by the forward half.  Thread = one 8x8 block, CTA = 6 warps x 32 macroblocks, block in a private shared-memory tile
// (layout of mb_encode_kernel).  A macroblock the stream did not transmit (MBA skipped it, or its whole GOB is missing)
// keeps the previous picture's samples: the decoder writes into a persistent picture buffer (CopyIob2FS, p64.c:1046).
// ---------------------------------------------------------------------------------------------------
struct MbDecArgs {
  Geom g;
  const uint8_t* ref;      // [S][frame_bytes] previous decoded picture (CFS)
  uint8_t* out;            // [S][frame_bytes] picture being decoded
  const p64b_mb* mbs;      // [S][nmb] GOB-major; reserved bit 0 = the macroblock was transmitted
  const int8_t* levels;    // [S][nmb][6][64] transmission order (intra DC as uint8)
  int n_streams;
};

template <int I>
__device__ __forceinline__ int level_at(const uint32_t (&w)[16]) {      // raster coefficient I = byte c_zig_at(I) (IZigzagMatrix, transform.c:546-553)
  constexpr int pos = c_zig_at(I);
  return (int)(int8_t)(w[pos >> 2] >> (8 * (pos & 3)));
}
template <int R>
__device__ __forceinline__ void dequant_row(int* tile, const uint32_t (&w)[16], int q2, int qo) {
  int v[8] = {level_at<8 * R + 0>(w), level_at<8 * R + 1>(w), level_at<8 * R + 2>(w), level_at<8 * R + 3>(w),
              level_at<8 * R + 4>(w), level_at<8 * R + 5>(w), level_at<8 * R + 6>(w), level_at<8 * R + 7>(w)};
#pragma unroll
  for (int j = 0; j < 8; j++) {                     // ICCITT[Flat]Quantize (transform.c:359-451) with the column pass's <<2 folded in:
    const int l = v[j];                             // 4 [(2|l|+1) Q - ev] sign(l) = l * 8Q + sign(l) * 4 (Q - ev); sign(0) = 0
    v[j] = l * q2 + min(max(l, -1), 1) * qo;
  }
  const int a[4] = {v[0], v[1], v[2], v[3]}, b[4] = {v[4], v[5], v[6], v[7]};
  st_row(tile, R, 0, a); st_row(tile, R, 1, b);
}

__global__ void __launch_bounds__(MB4_THREADS, 3)
mb_decode_kernel(const __grid_constant__ MbDecArgs a) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const Geom& g = a.g;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  int* tile = reinterpret_cast<int*>(s_dyn) + threadIdx.x * MB4_TILE;
  const int n_total = a.n_streams * g.nmb;
  const int n = blockIdx.x * MB4_PER_CTA + lane;
  if (n >= n_total) return;                          // no CTA-wide barrier below
  const int s = n / g.nmb, mbi = n - s * g.nmb;
  const int gob = mbi / 33, m = mbi - gob * 33;
  int col, row;                                      // MoveTo, io.c:730-741
  if (g.qcif) { col = m % 11; row = gob * 3 + m / 11; }
  else { col = (gob & 1) * 11 + m % 11; row = (gob >> 1) * 3 + m / 11; }
  const bool chroma = c >= 4;
  const int w = chroma ? g.W >> 1 : g.W, wq = w >> 3;
  const int off = chroma ? g.W * g.H + (c - 4) * (g.W * g.H >> 2) + row * 8 * w + col * 8
                         : (row * 16 + (c >> 1) * 8) * w + col * 16 + (c & 1) * 8;
  const size_t fo = (size_t)s * g.frame_bytes;
  uint2* op = reinterpret_cast<uint2*>(a.out + fo + off);
  const uint2 rw = __ldg(reinterpret_cast<const uint2*>(a.mbs + n));
  const int mt = rw.x & 0xff, cbp_raw = (rw.x >> 8) & 0xff, q = rw.y & 0xff;
  const int mvx = (int)(int8_t)(rw.x >> 16), mvy = (int)(int8_t)(rw.x >> 24);
  const bool present = (rw.y >> 16) & 1u;
  if (!present) {                                    // not transmitted: the picture buffer keeps the previous samples
    const uint2* rp = reinterpret_cast<const uint2*>(a.ref + fo + off);
#pragma unroll
    for (int r = 0; r < 8; r++) op[r * wq] = __ldg(rp + r * wq);
    return;
  }
  const bool intra = mt_is(M_INTRA, mt);
  const int cbp = mt_is(M_CBP, mt) ? cbp_raw : 0x3f;                       // p64.c:1196
  const bool coded = ((cbp >> (5 - c)) & 1) && mt_is(M_TCOEF, mt);

  // ---- prediction as packed bytes (Add*Compensate addressing, io.c:142-313; chroma vector = MV/2 truncating)
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; i++) pk[i] = 0;
  if (!intra) {
    int dx = 0, dy = 0;
    if (mt_is(M_MF, mt)) {
      dx = chroma ? mvx / 2 : mvx; dy = chroma ? mvy / 2 : mvy;
      // records come from a bit stream (untrusted): the block is kept inside the plane.  A conforming stream never needs it
      // (the parser rejects such vectors, decoder.cpp), so nothing the reference decodes changes.
      const int bx = chroma ? col * 8 : col * 16 + (c & 1) * 8, by = chroma ? row * 8 : row * 16 + (c >> 1) * 8;
      const int hh = chroma ? g.H >> 1 : g.H;
      dx = min(max(dx, -bx), w - 8 - bx); dy = min(max(dy, -by), hh - 8 - by);
    }
    const uint8_t* b = a.ref + fo + off + dy * w + dx;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint2 r = fetch_row8(b + i * w); pk[2 * i] = r.x; pk[2 * i + 1] = r.y; }
    if (mt_is(M_FILTER, mt)) {                       // LoadFilterMatrix, io.c:323-372 (filter_row: single rounding (S16+8)>>4)
      uint32_t f[16];
      filter_row<0, 4, 0>(pk, 0, 0, 0, f[0], f[1]);
      filter_row<1, 2, 1>(pk, 0, 1, 2, f[2], f[3]);   filter_row<1, 2, 1>(pk, 1, 2, 3, f[4], f[5]);
      filter_row<1, 2, 1>(pk, 2, 3, 4, f[6], f[7]);   filter_row<1, 2, 1>(pk, 3, 4, 5, f[8], f[9]);
      filter_row<1, 2, 1>(pk, 4, 5, 6, f[10], f[11]); filter_row<1, 2, 1>(pk, 5, 6, 7, f[12], f[13]);
      filter_row<0, 4, 0>(pk, 7, 7, 7, f[14], f[15]);
#pragma unroll
      for (int i = 0; i < 16; i++) pk[i] = f[i];
    }
  }
  if (!coded) {                                      // reconstruction = prediction (zero residual)
#pragma unroll
    for (int r = 0; r < 8; r++) op[r * wq] = make_uint2(pk[2 * r], pk[2 * r + 1]);
    return;
  }
  // ---- levels -> raster order -> inverse quantise -> tile
  {
    uint32_t lw[16];
    const uint4* lp = reinterpret_cast<const uint4*>(a.levels + ((size_t)n * 6 + c) * 64);
#pragma unroll
    for (int i = 0; i < 4; i++) { const uint4 v = __ldg(lp + i); lw[4 * i] = v.x; lw[4 * i + 1] = v.y; lw[4 * i + 2] = v.z; lw[4 * i + 3] = v.w; }
    const int ev = (q & 1) ? 0 : 1, q2 = 8 * q, qo = 4 * (q - ev);
    dequant_row<0>(tile, lw, q2, qo); dequant_row<1>(tile, lw, q2, qo); dequant_row<2>(tile, lw, q2, qo); dequant_row<3>(tile, lw, q2, qo);
    dequant_row<4>(tile, lw, q2, qo); dequant_row<5>(tile, lw, q2, qo); dequant_row<6>(tile, lw, q2, qo); dequant_row<7>(tile, lw, q2, qo);
    if (intra) tile[0] = (int)(lw[0] & 0xffu) * 32;                        // intra DC = 8 l (transform.c:362), x4 for the column pass
  }
  // ---- Chen IDCT: column pass on half blocks, then row pass + rounding + prediction + clamp, row by row
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    int v[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++) ld_row(tile, r, h, v[r]);
#pragma unroll
    for (int j = 0; j < 4; j++) idct8<2>(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
    for (int r = 0; r < 8; r++) st_row(tile, r, h, v[r]);
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    int x[8];
    {
      int p0[4], p1[4];
      ld_row(tile, r, 0, p0); ld_row(tile, r, 1, p1);
      x[0] = p0[0]; x[1] = p0[1]; x[2] = p0[2]; x[3] = p0[3]; x[4] = p1[0]; x[5] = p1[1]; x[6] = p1[2]; x[7] = p1[3];
    }
    idct8<0>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
    int o[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      o[j] = dp4a_us(pk[2 * r], 1 << (8 * j), round_div16(x[j]));
      o[4 + j] = dp4a_us(pk[2 * r + 1], 1 << (8 * j), round_div16(x[4 + j]));
    }
    op[r * wq] = make_uint2(pack4_sat(o[0], o[1], o[2], o[3]), pack4_sat(o[4], o[5], o[6], o[7]));
  }
}

// MBs overridden by the host to MType 4 with a zero vector when the rate buffer overflowed (p64.c:776-783):
// reconstruction = copy of the reference MB (AddCompensate at (0,0), TCoeffMType[4]=0), LastIntra++.
// One warp per (stream, MB); exits immediately when the MB was not overridden.
__global__ void overflow_patch_kernel(Geom g, const uint8_t* __restrict__ ovf, const uint8_t* __restrict__ ref,
                                      uint8_t* __restrict__ out, const uint8_t* __restrict__ li_prev,
                                      uint8_t* __restrict__ li_new, int n_streams) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_streams * g.nmb) return;
  if (!ovf[wid]) return;
  const int s = wid / g.nmb, mbi = wid % g.nmb, gob = mbi / 33, m = mbi % 33;
  int col, row;
  if (g.qcif) { col = m % 11; row = gob * 3 + m / 11; }
  else { col = (gob & 1) * 11 + m % 11; row = (gob >> 1) * 3 + m / 11; }
  const size_t fo = (size_t)s * g.frame_bytes;
  {  // luma 16 rows x 16 B: lane -> (row = lane/2, half = lane%2)
    size_t o = fo + (size_t)(row * 16 + lane / 2) * g.W + col * 16 + (lane & 1) * 8;
    *reinterpret_cast<uint2*>(out + o) = *reinterpret_cast<const uint2*>(ref + o);
  }
  if (lane < 16) {  // chroma: 2 planes x 8 rows x 8 B
    int pl = lane >> 3, r = lane & 7;
    size_t o = fo + (size_t)g.W * g.H + (size_t)pl * (g.W * g.H / 4) + (size_t)(row * 8 + r) * (g.W / 2) + col * 8;
    *reinterpret_cast<uint2*>(out + o) = *reinterpret_cast<const uint2*>(ref + o);
  }
  if (lane == 0) li_new[wid] = (uint8_t)(li_prev[wid] + 1);
}

// Frame statistics (StatisticsMem, stat.c:73-130): per (stream, plane) the exact integer sums the reference
// accumulates in double -- sum of the source, sum of the reconstruction, sum of squared differences, sum of squared source
// samples -- and the 256-bin histogram of the reconstruction.  One CTA per 8 KB chunk of a plane: 16-byte loads, packed-byte
// dot products for the sums, per-warp shared-memory histograms, one atomic per value into the plane's record.  The host derives
// mean / MSE / SNR / MRSNR / PSNR / entropy from the sums exactly as stat.c does (p64b_stat_from_sums).
constexpr int STATS_CHUNK = 8192;                // bytes of a plane per CTA: 256 threads x 2 x 16 bytes
__global__ void __launch_bounds__(256) plane_stats_kernel(Geom g, const uint8_t* __restrict__ src, const uint8_t* __restrict__ rec,
                                                          p64b_plane_stats* __restrict__ out, int n_streams) {
  __shared__ uint32_t s_hist[8][256];             // one histogram per warp: fewer same-address collisions on smooth content
  __shared__ unsigned long long s_sum[4];
  const int wh = g.W * g.H, cq = wh >> 2;
  const int chunks_y = (wh + STATS_CHUNK - 1) / STATS_CHUNK, chunks_c = (cq + STATS_CHUNK - 1) / STATS_CHUNK;
  const int per_stream = chunks_y + 2 * chunks_c;
  const int s = blockIdx.x / per_stream;
  int r = blockIdx.x - s * per_stream;
  if (s >= n_streams) return;
  int pl = 0, plane_off = 0, plane_n = wh;
  if (r >= chunks_y) { r -= chunks_y; pl = 1 + r / chunks_c; r -= (pl - 1) * chunks_c; plane_off = wh + (pl - 1) * cq; plane_n = cq; }
  const int begin = r * STATS_CHUNK, end = min(begin + STATS_CHUNK, plane_n);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
  if (threadIdx.x < 4) s_sum[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)s * g.frame_bytes + plane_off;     // frame_bytes, W*H and W*H/4 are multiples of 16
  uint32_t a_src = 0, a_rec = 0, a_err = 0, a_sq = 0;
  uint4 sv[2], rv[2];
  int idx[2];
#pragma unroll
  for (int u = 0; u < 2; u++) {                    // both loads of both planes in flight before any use
    idx[u] = begin + 16 * (threadIdx.x + 256 * u);
    sv[u] = rv[u] = make_uint4(0, 0, 0, 0);
    if (idx[u] < end) {
      sv[u] = __ldg(reinterpret_cast<const uint4*>(src + base + idx[u]));
      rv[u] = __ldg(reinterpret_cast<const uint4*>(rec + base + idx[u]));
    }
  }
#pragma unroll
  for (int u = 0; u < 2; u++) {
    if (idx[u] >= end) continue;
    const uint32_t sw[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w}, rw[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      a_src = __dp4a(sw[k], 0x01010101u, a_src);
      a_rec = __dp4a(rw[k], 0x01010101u, a_rec);
      a_sq = __dp4a(sw[k], sw[k], a_sq);
      a_err += __dp4a(sw[k], sw[k], 0u) + __dp4a(rw[k], rw[k], 0u) - 2u * __dp4a(sw[k], rw[k], 0u);
#pragma unroll
      for (int j = 0; j < 4; j++) atomicAdd(&s_hist[warp][(rw[k] >> (8 * j)) & 0xffu], 1u);
    }
  }
  a_src = __reduce_add_sync(0xffffffffu, a_src); a_rec = __reduce_add_sync(0xffffffffu, a_rec);
  a_err = __reduce_add_sync(0xffffffffu, a_err); a_sq = __reduce_add_sync(0xffffffffu, a_sq);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_sum[0], (unsigned long long)a_src); atomicAdd(&s_sum[1], (unsigned long long)a_rec);
    atomicAdd(&s_sum[2], (unsigned long long)a_err); atomicAdd(&s_sum[3], (unsigned long long)a_sq);
  }
  __syncthreads();
  p64b_plane_stats* o = out + (size_t)s * 3 + pl;
  uint32_t hsum = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) hsum += s_hist[w][threadIdx.x];
  if (hsum) atomicAdd(&o->hist[threadIdx.x], hsum);
  if (threadIdx.x == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(&o->sum_src), s_sum[0]);
    atomicAdd(reinterpret_cast<unsigned long long*>(&o->sum_rec), s_sum[1]);
    atomicAdd(reinterpret_cast<unsigned long long*>(&o->sum_sq_err), s_sum[2]);
    atomicAdd(reinterpret_cast<unsigned long long*>(&o->sum_sq_src), s_sum[3]);
    if (r == 0) o->n = (uint64_t)plane_n;
  }
}

// Register-only issue-rate probe for VABSDIFF4.U8.ACC (ME roofline denominator)
__global__ void __launch_bounds__(256) sad_peak_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[8], acc[8];
  uint32_t b = seed * 2654435761u + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = b * (i + 3) + blockIdx.x; acc[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(acc[i]) : "r"(x[i]), "r"(b));
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace p64b
