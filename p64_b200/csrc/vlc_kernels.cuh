// Device-side entropy coding: H.261 headers and run-level VLC on the GPU, so that a fixed-quantiser
// frame step returns the finished bit stream instead of records + levels.  Bit-exact restatement of the host coder
// (bits.cpp), i.e. of the reference's
//   WriteGOBHeader  marker.c:182-209      WriteMBHeader  marker.c:288-354 (MVD prediction 310-338)
//   EncodeDC / EncodeAC / CBPEncodeAC  codec.c:96-205, 346-355      WritePictureHeader  marker.c:103-137
//   mputv  stream.c:193-205 (MSB first)
//
//   vlc_gob_seq_kernel  fixed quantiser: a GOB is nine independent pieces of 4, 4 and 3 macroblocks; the warps of a resident
//                     grid take pieces from a queue and code one in one pass: one item (header half, intra DC, level,
//                     EOB) per lane, the words streamed out through a small ring in shared memory -- no measuring pass.
//   vlc_gob_kernel    rate control: one CTA per (stream, GOB): one thread per piece (GOB header, 33 x {MB header, 6 blocks});
//                     pass 1 measures every piece, a CTA-wide scan places them, pass 2 writes the bits into a
//                     shared-memory image of the GOB, which is then stored unshifted into the GOB's scratch slot.
//   vlc_sizes_kernel  whole bytes of every stream's chunk and their packed offsets.
//   vlc_frame_kernel  one CTA per stream: carry bits of the previous frame + picture header + the pieces (or whole GOBs)
//                     are concatenated into the frame's byte chunk: a word inside one piece is a funnel shift with the
//                     piece's own shift (streamed by a warp), a word that holds a piece boundary is assembled bit by
//                     bit; the < 8 trailing bits stay on the device as the next frame's carry, so the host only ever
//                     appends whole bytes.
// Bit strings are MSB first: bit i of a string is bit (31 - i % 32) of word i / 32.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/p64_b200.h"
#include "vlc_dev.h"

namespace p64b {

constexpr int VLC_PIECES = 1 + 33 * 7;                  // GOB header + 33 x (MB header + 6 blocks)
constexpr int VLC_THREADS = 256;
constexpr int VLC_BLOCK_MAX_BITS = 8 + 63 * 20 + 2;     // intra DC + 63 escapes + EOB
constexpr int VLC_MBHDR_MAX_BITS = 1 + 10 + 5 + 11 + 11 + 9;
// The fixed-quantiser kernel writes a GOB as VLC_PPG independent pieces of 4, 4 and 3 macroblocks per macroblock row; the frame
// kernel concatenates pieces exactly as it concatenates whole GOBs (rate control: one piece per GOB).  Piece `pc` starts at
// macroblock vlc_piece_m0(pc) of the GOB and at word m0 * VLC_MB_WORDS of the GOB's slot.
constexpr int VLC_PPG = 9;
__host__ __device__ constexpr int vlc_piece_m0(int pc) { return 11 * (pc / 3) + 4 * (pc % 3); }
__host__ __device__ constexpr int vlc_piece_n(int pc) { return pc % 3 == 2 ? 3 : 4; }
constexpr int VLC_MB_WORDS = (VLC_MBHDR_MAX_BITS + 6 * VLC_BLOCK_MAX_BITS) / 32 + 4;   // n >= 1 macroblocks + the GOB header + a partial and a look-ahead word fit n of these
static_assert((26 + (VLC_MBHDR_MAX_BITS + 6 * VLC_BLOCK_MAX_BITS) + 31) / 32 + 1 <= VLC_MB_WORDS, "a piece fits its part of the slot");
constexpr int VLC_GOB_WORDS_1 = (26 + 33 * (VLC_MBHDR_MAX_BITS + 6 * VLC_BLOCK_MAX_BITS) + 31) / 32 + 8;
constexpr int VLC_GOB_WORDS = VLC_GOB_WORDS_1 > 33 * VLC_MB_WORDS ? VLC_GOB_WORDS_1 : 33 * VLC_MB_WORDS;   // slot size, words
constexpr int VLC_PIC_HDR_MAX_BITS = 41;

// Rate control on the device ("exact per-MB bit counts on device let GQUANT selection and the overflow
// test run without a per-GOB host round trip").  Restates, per stream,
//   BufferContents()  p64.c:233-236   mwtell() + BufferOffset - ((CurrentGOB*33 + CurrentMDU) * Rate*FrameSkip / denom)
//   BufferSize()      p64.c:237       Rate / 4
//   ExecuteQuantization p64.c:458-481 GQuant = clamp(BufferContents()/QDFact + QOffs, 1, 31), every GOB but not on the
//                                     first frame (p64.c:697-702)
//   the per-macroblock overflow override p64.c:776-783 (MType 4, zero vector; also on the first frame)
//   the end-of-frame BufferOffset update p64.c:670-680
// in the reference's integer types (int, with the long->int truncations where the reference has them).
struct RcArgs {
  int rate;                        // -r bits/s; 0 = rate control off
  int frame_skip, frame_rate, frame_rate_div, qdfact, qoffs;
  int denom;                       // NumberGOB*NumberMDU*FrameRate/FrameRateDiv (p64.c:236)
  int first_frame;
  int pic_hdr_bits;
  const unsigned long long* bitpos;  // [S] mwtell() at the end of the previous frame
  uint32_t* frame_bits;            // [S] bits of this frame so far: picture header + the GOBs already coded
  long long* buffer_offset;        // [S] BufferOffset (p64.c:136)
  uint8_t* quant;                  // [S] GQuant: read for this GOB, written for the next one
  uint8_t* ovf;                    // [S][nmb] macroblocks overridden in this frame (for the reconstruction patch)
  uint32_t* overflows;             // [S] NumberOvfl, cumulative (p64.c:780)
};
__host__ __device__ inline long long rc_buffer_contents(const RcArgs& r, long long tell, long long offset, int g, int m) {
  const int num = (int)((long long)(g * 33 + m) * r.rate * r.frame_skip);
  return tell + offset - (r.denom ? num / r.denom : 0);
}
__host__ __device__ inline int rc_execute_quantization(const RcArgs& r, long long tell, long long offset, int g) {
  const int cur = (int)rc_buffer_contents(r, tell, offset, g, 0);
  const int q = cur / r.qdfact + r.qoffs;
  return q < 1 ? 1 : (q > 31 ? 31 : q);
}

struct VlcArgs {
  const DevVlcTables* tables;
  const p64b_mb* mbs;        // [S][nmb] GOB-major
  const int8_t* levels;      // [S][nmb][6][64] transmission order
  uint32_t* gob_words;       // [S][ngob][VLC_GOB_WORDS] scratch; the fixed-quantiser kernel puts piece pc at vlc_piece_m0(pc) * VLC_MB_WORDS of the slot
  uint32_t* gob_bits;        // [S][ngob] (rate control) or [S][ngob][VLC_PPG]
  uint32_t* queue;           // [2] piece counters of the fixed-quantiser kernel: this launch takes from queue[parity], zeroes the other
  int parity;
  int n_streams, ngob, nmb;
  int qcif;                  // GOB numbers 1,3,5 (p64.c:709-711)
  int gquant;                // GQUANT of every stream when rc.rate == 0
  int gob_first, gob_count;  // CTA b -> stream b / gob_count, GOB gob_first + b % gob_count
  RcArgs rc;
};

// MType property tables (p64.c:217-222) as bit masks over type 0..9
constexpr uint32_t V_QUANT = 0x24a, V_CBP = 0x36c, V_MF = 0x3f0, V_TCOEF = 0x36f;
__device__ __forceinline__ bool vt(uint32_t mask, int mt) { return (mask >> mt) & 1u; }

// Writes a piece's bits at its final position in a zeroed word buffer.  Words that lie completely inside the piece
// are owned by it (plain stores); the partial first and last words are shared with the neighbours (atomic OR).
struct BitEmitter {
  uint32_t* buf;
  uint64_t acc = 0;      // pending bits, right-aligned
  int cnt = 0;           // number of pending bits (< 32 between puts)
  int word;              // next word to write
  int room;              // bits still free in that word
  __device__ __forceinline__ BitEmitter(uint32_t* b, uint32_t bit_offset) : buf(b), word((int)(bit_offset >> 5)), room(32 - (int)(bit_offset & 31)) {}
  __device__ __forceinline__ void put(uint32_t v, int n) {       // n <= 32, v < 2^n
    acc = (acc << n) | v;
    cnt += n;
    while (cnt >= room) {
      const uint32_t out = (uint32_t)(acc >> (cnt - room)) & (room == 32 ? 0xffffffffu : ((1u << room) - 1u));
      if (room == 32) buf[word] = out; else atomicOr(buf + word, out);
      cnt -= room; word++; room = 32;
      acc &= (1ull << cnt) - 1ull;
    }
  }
  __device__ __forceinline__ void flush() {
    if (cnt) atomicOr(buf + word, (uint32_t)(acc << (room - cnt)));
  }
};

// One 8x8 block (EncodeDC + EncodeAC for intra types, CBPEncodeAC otherwise; codec.c:96-205, 346-355).
// EMIT = false: returns the length in bits; EMIT = true: also writes the bits.
// non-zero mask of a block's 64 levels (bit k = level k != 0)
__device__ __forceinline__ uint64_t vlc_nz_mask(const int8_t* __restrict__ lv) {
  uint64_t nz = 0;
  const uint4* lp = reinterpret_cast<const uint4*>(lv);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint4 q = __ldg(lp + i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t hi = (w[j] | ((w[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;     // bit 7 of every non-zero byte
      const uint32_t m4 = (((hi >> 7) * 0x01020408u) >> 24) & 0xfu;
      nz |= (uint64_t)m4 << (16 * i + 4 * j);
    }
  }
  return nz;
}

// `nz`: the block's non-zero mask (vlc_nz_mask), computed once by the caller and reused by the emitting pass
template <bool EMIT>
__device__ __forceinline__ int vlc_block(const int8_t* __restrict__ lv, uint64_t nz, bool cbp_type, const uint32_t* s_tcoef, BitEmitter* em) {
  int len = 0, prev = -1;
  bool first = cbp_type, any = !cbp_type;
  if (!cbp_type) {                                   // EncodeDC, codec.c:346-355
    int dc = (uint8_t)lv[0];
    dc = min(max(dc, 1), 254);
    if (dc == 128) dc = 255;
    if (EMIT) em->put((uint32_t)dc, 8);
    len = 8; prev = 0; nz &= ~1ull;
  }
  while (nz) {
    const int p = __ffsll((long long)nz) - 1;
    nz &= nz - 1;
    const int run = p - prev - 1, v = lv[p], a = abs(v);
    prev = p;
    if (first && run == 0 && a == 1) {               // "1s" for the first coefficient of a CBP-coded block
      if (EMIT) em->put(2u | (uint32_t)(v < 0), 2);
      len += 2;
    } else {
      const uint32_t e = (run < 32 && a < 16) ? s_tcoef[run * 16 + a] : 0u;
      if (e) {
        const int n = (int)(e >> 16) + 1;
        if (EMIT) em->put(((e & 0xffffu) << 1) | (uint32_t)(v < 0), n);
        len += n;
      } else {                                       // escape: 000001 + 6-bit run + 8-bit level (codec.c:113-115)
        if (EMIT) em->put((1u << 14) | ((uint32_t)run << 8) | (uint32_t)(v & 0xff), 20);
        len += 20;
      }
    }
    first = false; any = true;
  }
  if (any) {                                         // EOB "10"; an all-zero CBP block gets none (codec.c:169-174)
    if (EMIT) em->put(2u, 2);
    len += 2;
  }
  return len;
}

// Macroblock header (WriteMBHeader, marker.c:288-354): MBA (always 1: this encoder never skips a macroblock,
// p64.c:928-930), MTYPE, [MQUANT], [MVD pair], [CBP].  Returns the bits right-aligned, *n = length (<= 47).
__device__ __forceinline__ uint64_t vlc_mb_header(const p64b_mb& r, const p64b_mb& prev, int m, const DevVlcTables* t, int* n) {
  uint64_t b = 1; int len = 1;
  auto put = [&](uint32_t e) { b = (b << (e >> 16)) | (e & 0xffffu); len += (int)(e >> 16); };
  const int mt = r.mtype;
  put(t->mtype[mt]);
  if (vt(V_QUANT, mt)) { b = (b << 5) | r.quant; len += 5; }
  if (vt(V_MF, mt)) {                                 // marker.c:310-338
    int h = r.mvx, v = r.mvy;
    if (m != 0 && m != 11 && m != 22 && vt(V_MF, prev.mtype)) {
      h -= prev.mvx; v -= prev.mvy;
      if (h < -16) h += 32;
      if (h > 15) h -= 32;
      if (v < -16) v += 32;
      if (v > 15) v -= 32;
    }
    put(t->mvd[h & 31]); put(t->mvd[v & 31]);
  }
  if (vt(V_CBP, mt)) put(t->cbp[r.cbp]);
  *n = len;
  return b;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fixed quantiser (round 2): a GOB is written as VLC_PPG independent PIECES of 4, 4 and 3 macroblocks per macroblock row
// (a macroblock's bits depend on nothing but its own record and its left neighbour's vector: MBA is always 1, the
// quantiser is fixed); ONE WARP codes a piece in ONE pass, one ITEM per lane.
// A piece's bit string is, per macroblock, the header followed by items in a fixed order: per coded block [intra DC], the
// non-zero levels in transmission order, [EOB].  The header counts as two more items (its bits above and below bit 32).
//   list   lane 6j+b (macroblock j of the piece, block b) holds the block's 64 levels: it stages them in shared memory
//          and writes the descriptors of its items (lane | kind | position) into the warp's list at the position a warp
//          scan of the counts gives: stream order; lanes 24+j build the headers (WriteMBHeader, marker.c:288-354);
//   code   lane k takes item k, 32 at a time: the run comes from the previous descriptor, the code from the table in shared
//          memory (codec.c:96-205: escapes as in 113-115, the two-bit code for a leading +-1 of a CBP-coded block; EncodeDC
//          codec.c:346-355); a warp scan of the lengths places the items;
//   emit   every lane ORs its item into a 64-word ring in shared memory; the words the position has passed are stored to the
//          piece's scratch slot and zeroed.
// The warps of a resident grid take pieces from a queue.  Every level is looked at once, nothing is measured first and
// placed later, no CTA-wide barrier after the table load.
// History (256 CIF frames per launch): thread per (GOB header | MB header | block), measure / scan / write: 71 M
// warp-instructions, 0.116 ms (round 1; still the rate-control kernel, one GOB per stream per launch).  One warp per GOB,
// macroblock after macroblock, lane per block quarter for the list: 57.9 M, 0.108 ms at 21 warps per SM (3 072 GOBs).  The same
// per piece of 3 macroblocks from a queue: 61 M, 0.090 ms.  This kernel: the list, the headers and the scans are shared by
// 4 macroblocks, the item rounds are full: 42 M, 0.073 ms (profiles/r02_ncu_vlc_pieces_final.txt).
constexpr int VLC_SEQ_WARPS = 4;
constexpr int VLC_SEQ_THREADS = 32 * VLC_SEQ_WARPS;
constexpr int VLC_PASS_MBS = 4;
constexpr int VLC_ITEMS_MAX = VLC_PASS_MBS * (6 * 65 + 2);   // per block: DC or position 0, 63 levels, EOB; + the header halves
constexpr int VLC_RING = 64;                            // words; one round adds at most 4 x 47 + 32 x 20 bits = 26 words

__device__ __forceinline__ uint32_t vlc_nz16(uint4 v) {      // bit k = byte k of the 16 is non-zero
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t nz = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t hi = (w[j] | ((w[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;     // bit 7 of every non-zero byte
    nz |= ((((hi >> 7) * 0x01020408u) >> 24) & 0xfu) << (4 * j);
  }
  return nz;
}
// OR `n` bits (n <= 32, v < 2^n) into the ring of VLC_RING zeroed words that holds the MSB-first string around the ABSOLUTE
// bit offset `o`
__device__ __forceinline__ void vlc_or_ring(uint32_t* ring, uint32_t o, uint32_t v, int n) {
  const uint32_t w = o >> 5;
  const uint64_t win = (uint64_t)v << ((64 - (int)(o & 31) - n) & 63);     // the bits in the 64-bit window that starts at word w (n = 0: v = 0)
  const uint32_t hi = (uint32_t)(win >> 32), lo = (uint32_t)win;
  if (hi) atomicOr(ring + (w & (VLC_RING - 1)), hi);
  if (lo) atomicOr(ring + ((w + 1) & (VLC_RING - 1)), lo);
}

// next piece of the launch's queue, taken by lane 0 (the same predicated per-lane-address atomic as the motion search's queue:
// a plain atomicAdd is warp-aggregated and its result broadcast waits for the atomic at once)
__device__ __forceinline__ uint32_t vlc_queue_take(uint32_t* counter, int lane) {
  uint32_t t = 0;
  asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %2, 0;\n@p atom.global.add.u32 %0, [%1], 1;\n}" : "+r"(t) : "l"(counter + lane), "r"(lane) : "memory");
  return t;
}

#ifndef P64B_VLC_MIN_CTAS
#define P64B_VLC_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(VLC_SEQ_THREADS, P64B_VLC_MIN_CTAS)
vlc_gob_seq_kernel(const __grid_constant__ VlcArgs a) {
  __shared__ DevVlcTables s_t;
  __shared__ __align__(16) uint8_t s_stage[VLC_SEQ_WARPS][24 * 64];
  __shared__ uint16_t s_items[VLC_SEQ_WARPS][VLC_ITEMS_MAX + 8];
  __shared__ uint32_t s_ring[VLC_SEQ_WARPS][VLC_RING];
  __shared__ uint2 s_lit[VLC_SEQ_WARPS][24 * 4];      // (bits, length) of the items that are not levels: per block-lane DC, header above bit 32, header low 32 bits, EOB
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 0 && tid == 0) a.queue[a.parity ^ 1] = 0;     // the next launch's counter
  for (int i = tid; i < DEV_VLC_WORDS; i += VLC_SEQ_THREADS) reinterpret_cast<uint32_t*>(&s_t)[i] = reinterpret_cast<const uint32_t*>(a.tables)[i];
  for (int i = lane; i < VLC_RING; i += 32) s_ring[warp][i] = 0;
  __syncthreads();                                      // (no barrier below)
  const int n_tasks = a.n_streams * a.ngob * VLC_PPG, n_workers = (int)gridDim.x * VLC_SEQ_WARPS;
  uint8_t* stage = s_stage[warp];
  uint16_t* items = s_items[warp];
  uint32_t* ring = s_ring[warp];
  uint2* lit = s_lit[warp];
  const int j = lane < 24 ? lane / 6 : lane - 24, b = lane - 6 * j;       // lanes 0..23: block b of macroblock j; 24..27: header of j
  int task = blockIdx.x * VLC_SEQ_WARPS + warp;         // first piece: static; the rest come from the queue
  uint32_t ticket = vlc_queue_take(a.queue + a.parity, lane);
#pragma unroll 1
  while (task < n_tasks) {
    const int task_next = n_workers + (int)__shfl_sync(0xffffffffu, ticket, 0);
    if (task_next < n_tasks) ticket = vlc_queue_take(a.queue + a.parity, lane);
    // task = (stream * ngob + gob) * VLC_PPG + piece (this kernel always codes whole frames: nmb = 33 ngob)
    const int pc = task % VLC_PPG, sg = task / VLC_PPG;
    const int m0 = vlc_piece_m0(pc), nm = vlc_piece_n(pc);
    const size_t mb0 = (size_t)sg * 33 + m0;
    uint32_t* dst = a.gob_words + (size_t)sg * VLC_GOB_WORDS + m0 * VLC_MB_WORDS;

    // ---- loads: the macroblock's record; lanes 0..23 their block's levels; lanes 24..27 the left neighbour's record
    const bool mine = j < nm;
    p64b_mb rec{}, prev{};
    uint4 lv[4] = {};
    if (mine) {
      const uint2 rw = __ldg(reinterpret_cast<const uint2*>(a.mbs + mb0 + j));
      rec = *reinterpret_cast<const p64b_mb*>(&rw);
      if (lane >= 24) {
        if (m0 + j) { const uint2 pw = __ldg(reinterpret_cast<const uint2*>(a.mbs + mb0 + j - 1)); prev = *reinterpret_cast<const p64b_mb*>(&pw); }
      } else {
        const uint4* lp = reinterpret_cast<const uint4*>(a.levels + (mb0 + j) * P64B_LEVELS_PER_MB + b * 64);
#pragma unroll
        for (int i = 0; i < 4; i++) lv[i] = __ldg(lp + i);
      }
    }
    uint32_t pos = 0, wbase = 0;                         // bits written so far; first word of the piece's string still in the ring
    if (m0 == 0) {                                       // WriteGOBHeader, marker.c:182-209: GBSC, GN, GQUANT, no GSPARE
      const int gob = sg % a.ngob;
      const int gn = (a.qcif ? (gob << 1) : gob) + 1;
      if (lane == 0) vlc_or_ring(ring, 0, (1u << 10) | ((uint32_t)gn << 6) | ((uint32_t)a.gquant << 1), 26);
      pos = 26;
    }
    // the words [wbase, pos / 32) are final: store them and hand their ring slots back
    auto flush = [&]() {
      __syncwarp();
      const uint32_t n = (pos >> 5) - wbase;              // <= 32
      if ((uint32_t)lane < n) { const uint32_t w = wbase + (uint32_t)lane; dst[w] = ring[w & (VLC_RING - 1)]; ring[w & (VLC_RING - 1)] = 0; }
      wbase += n;
      __syncwarp();
    };
    // ---- list
    const int mt = rec.mtype;
    const bool cbp_type = vt(V_CBP, mt);
    const bool coded = mine && lane < 24 && vt(V_TCOEF, mt) && ((rec.cbp >> (5 - b)) & 1);
    uint32_t nzl = 0, nzh = 0;
    if (lane >= 24) {
      if (mine) {
        int hl;
        const uint64_t hb = vlc_mb_header(rec, prev, m0 + j, &s_t, &hl);
        lit[24 * j + 1] = make_uint2((uint32_t)(hb >> 32), (uint32_t)max(hl - 32, 0));
        lit[24 * j + 2] = make_uint2((uint32_t)hb, (uint32_t)min(hl, 32));
      }
    } else {
      if (!coded) { lv[0] = lv[1] = lv[2] = lv[3] = make_uint4(0, 0, 0, 0); }      // (the levels were loaded before the record was known)
#pragma unroll
      for (int i = 0; i < 4; i++) *reinterpret_cast<uint4*>(stage + lane * 64 + i * 16) = lv[i];
      nzl = vlc_nz16(lv[0]) | (vlc_nz16(lv[1]) << 16);
      nzh = vlc_nz16(lv[2]) | (vlc_nz16(lv[3]) << 16);
    }
    const bool dc = coded && !cbp_type;
    if (dc) {                                              // the intra DC is its own item: EncodeDC, codec.c:346-355
      nzl &= ~1u;
      int v = (int)(lv[0].x & 0xffu);
      v = min(max(v, 1), 254);
      if (v == 128) v = 255;
      lit[4 * lane] = make_uint2((uint32_t)v, 8u);
    }
    const bool eob = coded && (!cbp_type || (nzl | nzh));  // an all-zero CBP block gets no EOB (codec.c:169-174)
    if (eob) lit[4 * lane + 3] = make_uint2(2u, 2u);       // "10"
    const bool head = mine && lane < 24 && b == 0;         // the macroblock's header goes in front of its first block
    const int cnt = __popc(nzl) + __popc(nzh) + (dc ? 1 : 0) + (eob ? 1 : 0) + (head ? 2 : 0);
    int x = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    const int total = __shfl_sync(0xffffffffu, x, 31);
    {
      int idx = x - cnt;
      const uint32_t base = (uint32_t)lane << 8;
      if (head) { items[idx++] = (uint16_t)(base | 0xc0u); items[idx++] = (uint16_t)(base | 0xc1u); }
      if (dc) items[idx++] = (uint16_t)(base | 0x80u);
      // (bit-reversed masks: the next position is a count of leading zeros, one instruction)
      for (uint32_t r = __brev(nzl); r;) { const int pl = __clz((int)r); r &= ~(0x80000000u >> pl); items[idx++] = (uint16_t)(base | (uint32_t)pl); }
      for (uint32_t r = __brev(nzh); r;) { const int pl = __clz((int)r); r &= ~(0x80000000u >> pl); items[idx++] = (uint16_t)(base | 32u | (uint32_t)pl); }
      if (eob) items[idx++] = (uint16_t)(base | 0xc2u);
    }
    __syncwarp();
    // ---- code + emit, 32 items at a time
#pragma unroll 1
    for (int k0 = 0; k0 < total; k0 += 32) {
      const int k = k0 + lane;
      uint32_t code = 0; int len = 0;
      if (k < total) {
        // descriptors: a level = lane << 8 | position; the others have bit 7 set: DC 0x80 (as a predecessor it reads as
        // position 0), header halves 0xc0 / 0xc1, EOB 0xc2 -- their bits come from `lit`
        const uint32_t d = items[k];
        const int L = (int)(d >> 8);
        if (d & 0x80u) {
          const uint2 t = lit[4 * L + (int)(d & 3u) + (int)((d >> 6) & 1u)];
          code = t.x; len = (int)t.y;
        } else {
          const int p = (int)(d & 63u);
          const int v = (int)(int8_t)stage[L * 64 + p], av = abs(v);
          const uint32_t pd = items[k - 1];                // (item 0 is a header half: k > 0 here)
          const int prevp = ((int)(pd >> 8) == L && !(pd & 0x40u)) ? (int)(pd & 63u) : -1;     // the block's DC or previous level
          const int run = p - prevp - 1;
          if (p == 0 && av == 1) {                         // "1s": a leading +-1 of a CBP-coded block (only those list position 0)
            code = 2u | (uint32_t)(v < 0); len = 2;
          } else {
            const uint32_t e = (run < 32 && av < 16) ? s_t.tcoef[run * 16 + av] : 0u;
            if (e) { len = (int)(e >> 16) + 1; code = ((e & 0xffffu) << 1) | (uint32_t)(v < 0); }
            else { code = (1u << 14) | ((uint32_t)run << 8) | (uint32_t)(v & 0xff); len = 20; }     // escape (codec.c:113-115)
          }
        }
      }
      int y = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int z = __shfl_up_sync(0xffffffffu, y, d); if (lane >= d) y += z; }
      vlc_or_ring(ring, pos + (uint32_t)(y - len), code, len);
      pos += (uint32_t)__shfl_sync(0xffffffffu, y, 31);
      if ((pos >> 5) - wbase >= 7) flush();               // (a round adds at most 26 words: never more than 32 to store, 59 in the ring)
    }
    flush();
    const uint32_t nwords = (pos + 31) >> 5;               // the partial last word, + one zero word for the gather's look-ahead
    if (lane == 0) {
      const uint32_t w = (pos >> 5) & (VLC_RING - 1);
      if (pos & 31) { dst[pos >> 5] = ring[w]; ring[w] = 0; }    // (the ring is all zero again for the warp's next piece)
      dst[nwords] = 0;
      a.gob_bits[task] = pos;
    }
    __syncwarp();
    task = task_next;
  }
}

// exclusive scan of one value per thread over the CTA; *total = sum.  Two barriers.
__device__ __forceinline__ uint32_t vlc_cta_scan(uint32_t len, uint32_t* s_wsum, uint32_t* s_total, uint32_t* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t x = len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  __syncthreads();                                     // s_wsum / s_total of an earlier scan have been read
  if (lane == 31) s_wsum[warp] = x;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; w++) base += s_wsum[w];
  if (tid == VLC_THREADS - 1) *s_total = base + x;
  __syncthreads();
  *total = *s_total;
  return base + x - len;
}

// RC = true: one GOB of every stream under rate control.  Pieces are first measured as if no macroblock were
// overridden; the bit position of every macroblock header then gives the overflow test (p64.c:776) of all 33 macroblocks at
// once, exact up to and including the first one that fires.  Only when one fires does thread 0 walk the GOB in order
// (an overridden macroblock is MType 4 with a zero vector and no coefficients, which also changes the MVD predictor of
// its successor, marker.c:310-338), after which the pieces are re-placed.  The CTA then publishes the GOB's length and
// the next GOB's GQUANT (ExecuteQuantization, p64.c:458-481).
template <bool RC>
__global__ void __launch_bounds__(VLC_THREADS)
vlc_gob_kernel(const __grid_constant__ VlcArgs a) {
  __shared__ uint32_t s_buf[VLC_GOB_WORDS];
  __shared__ DevVlcTables s_t;
  __shared__ uint32_t s_total;
  __shared__ uint32_t s_wsum[VLC_THREADS / 32];
  __shared__ uint32_t s_len[RC ? VLC_THREADS : 1];
  __shared__ uint2 s_rec[RC ? 33 : 1];
  __shared__ unsigned long long s_hb[RC ? 33 : 1];
  __shared__ uint8_t s_hl[RC ? 33 : 1], s_ovf[RC ? 33 : 1];
  const int tid = threadIdx.x;
  const int s = blockIdx.x / a.gob_count, gob = a.gob_first + blockIdx.x % a.gob_count;
  for (int i = tid; i < DEV_VLC_WORDS; i += VLC_THREADS) reinterpret_cast<uint32_t*>(&s_t)[i] = reinterpret_cast<const uint32_t*>(a.tables)[i];
  const int gquant = RC ? (int)a.rc.quant[s] : a.gquant;
  long long tell0 = 0, boff = 0;                      // mwtell() before this GOB's header; BufferOffset
  if (RC) { tell0 = (long long)a.rc.bitpos[s] + a.rc.frame_bits[s]; boff = a.rc.buffer_offset[s]; }
  __syncthreads();

  // ---- this thread's piece
  const int piece = tid;                              // 0: GOB header; 1 + 7m: header of MB m; 1 + 7m + 1 + c: block c
  const bool is_piece = piece < VLC_PIECES;
  const int m = piece ? (piece - 1) / 7 : 0, k = piece ? (piece - 1) % 7 : 0;
  const size_t mbi = (size_t)s * a.nmb + gob * 33 + m;
  p64b_mb rec{}, prev{};
  uint64_t hbits = 0;
  int len = 0;
  bool coded = false;
  uint64_t nzmask = 0;
  const int8_t* lv = a.levels + (mbi * 6 + (k ? k - 1 : 0)) * 64;
  if (is_piece) {
    if (piece == 0) {                                 // WriteGOBHeader, marker.c:182-209: GBSC, GN, GQUANT, no GSPARE
      const int gn = (a.qcif ? (gob << 1) : gob) + 1;
      hbits = (1ull << 10) | ((uint64_t)gn << 6) | ((uint64_t)gquant << 1);
      len = 26;
    } else {
      rec = a.mbs[mbi];
      if (k == 0) {
        if (m) prev = a.mbs[mbi - 1];
        hbits = vlc_mb_header(rec, prev, m, &s_t, &len);
        if (RC) s_rec[m] = *reinterpret_cast<const uint2*>(&rec);
      } else {
        coded = vt(V_TCOEF, rec.mtype) && ((rec.cbp >> (6 - k)) & 1);      // block c = k-1: bit 5-c
        if (coded) { nzmask = vlc_nz_mask(lv); len = vlc_block<false>(lv, nzmask, vt(V_CBP, rec.mtype), s_t.tcoef, nullptr); }
      }
    }
  }

  // ---- placement: exclusive scan of the piece lengths
  uint32_t total;
  uint32_t off = vlc_cta_scan((uint32_t)len, s_wsum, &s_total, &total);
  if (RC) {
    const int bsize = a.rc.rate / 4;                   // BufferSize(), p64.c:237
    s_len[tid] = (uint32_t)len;
    const bool fires = is_piece && piece && k == 0 && rc_buffer_contents(a.rc, tell0 + off, boff, gob, m) > bsize;
    if (__syncthreads_or(fires)) {
      if (tid == 0) {
        uint32_t bits = 26;
        p64b_mb pe{};
        for (int i = 0; i < 33; i++) {
          p64b_mb r = *reinterpret_cast<const p64b_mb*>(&s_rec[i]);
          const bool o = rc_buffer_contents(a.rc, tell0 + bits, boff, gob, i) > bsize;
          if (o) { r = p64b_mb{}; r.mtype = 4; r.cbp = 0x3f; r.quant = (uint8_t)gquant; }      // p64.c:778-779
          int n;
          s_hb[i] = vlc_mb_header(r, pe, i, &s_t, &n);
          s_hl[i] = (uint8_t)n; s_ovf[i] = o;
          bits += (uint32_t)n;
          if (!o) for (int c = 0; c < 6; c++) bits += s_len[1 + 7 * i + 1 + c];
          pe = r;
        }
      }
      __syncthreads();
      if (is_piece && piece) {
        if (k == 0) { hbits = s_hb[m]; len = s_hl[m]; a.rc.ovf[mbi] = s_ovf[m]; }
        else if (s_ovf[m]) len = 0;
      }
      off = vlc_cta_scan((uint32_t)len, s_wsum, &s_total, &total);
      if (tid < 32) {                                  // NumberOvfl
        uint32_t n = (uint32_t)s_ovf[tid] + (tid == 0 ? (uint32_t)s_ovf[32] : 0u);
#pragma unroll
        for (int d = 16; d; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
        if (tid == 0) a.rc.overflows[s] += n;
      }
    }
  }
  const uint32_t nwords = (total + 31) >> 5;
  for (uint32_t i = tid; i < nwords + 1; i += VLC_THREADS) s_buf[i] = 0;
  __syncthreads();

  // ---- pass 2: the bits
  if (len) {
    BitEmitter em(s_buf, off);
    if (k == 0) {
      if (len > 32) { em.put((uint32_t)(hbits >> 32), len - 32); em.put((uint32_t)hbits, 32); }
      else em.put((uint32_t)hbits, len);
    } else {
      vlc_block<true>(lv, nzmask, vt(V_CBP, rec.mtype), s_t.tcoef, &em);
    }
    em.flush();
  }
  __syncthreads();
  uint32_t* dst = a.gob_words + ((size_t)s * a.ngob + gob) * VLC_GOB_WORDS;
  for (uint32_t i = tid; i < nwords + 1; i += VLC_THREADS) dst[i] = s_buf[i];      // + one zero word for the gather's look-ahead
  if (tid == 0) {
    a.gob_bits[(size_t)s * a.ngob + gob] = total;
    if (RC) {
      const uint32_t fb = a.rc.frame_bits[s] + total;
      a.rc.frame_bits[s] = fb;
      if (!a.rc.first_frame && gob + 1 < a.ngob)       // GQUANT of the next GOB (p64.c:697-702)
        a.rc.quant[s] = (uint8_t)rc_execute_quantization(a.rc, (long long)a.rc.bitpos[s] + fb, boff, gob + 1);
    }
  }
}

// the picture header of the step about to run (kernel parameters -> device memory; launched outside any graph)
__global__ void set_pic_hdr_kernel(uint32_t* dst, uint32_t w0, uint32_t w1) { dst[0] = w0; dst[1] = w1; }

// Start of a frame under rate control: the picture header is written before GOB 0 asks for its quantiser.
__global__ void rc_frame_begin_kernel(RcArgs r, int n_streams, int initial_quant) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  r.frame_bits[s] = (uint32_t)r.pic_hdr_bits;
  if (initial_quant) r.quant[s] = (uint8_t)initial_quant;                 // first frame of the sequence (p64.c:574-590)
  else if (!r.first_frame)
    r.quant[s] = (uint8_t)rc_execute_quantization(r, (long long)r.bitpos[s] + r.pic_hdr_bits, r.buffer_offset[s], 0);
}

// Output of one frame step, one buffer (copied to the host in one piece):
//   uint32 offset[S+1]   byte offset of stream s's chunk in data[] (multiples of 4); offset[S] = bytes used
//   uint32 nbytes[S]     whole bytes of the chunk
//   uint32 carry[S]      the stream's pending bits (< 8) after this frame, left-aligned in the word
//   uint32 carry_len[S]
//   uint64 bitpos[S]     bits written so far (mwtell, stream.c:233-238)
//   uint32 gquant[S]     GQuant after this frame (rate control; else the step's quantiser)
//   uint32 overflows[S]  NumberOvfl so far (rate control; else 0)
//   uint8  data[]        at byte vlc_data_offset(S)
__host__ __device__ constexpr size_t vlc_bitpos_offset(int S) { return (((size_t)(4 * S + 1) * 4 + 15) / 16) * 16; }
__host__ __device__ constexpr size_t vlc_data_offset(int S) { return vlc_bitpos_offset(S) + (size_t)S * 16; }
struct VlcFrameArgs {
  const uint32_t* gob_words;   // [S][ngob][VLC_GOB_WORDS]
  const uint32_t* gob_bits;    // [S][ngob]
  uint32_t* carry;             // [S] stream state, resident
  uint32_t* carry_len;         // [S]
  unsigned long long* bitpos;  // [S]
  uint8_t* out;                // the step's output buffer (layout above)
  int n_streams, ngob;
  int ppg = 1;                 // pieces per GOB: 1 with rate control, VLC_PPG without (at vlc_piece_m0() * VLC_MB_WORDS of the slot)
  const uint32_t* pic_hdr;     // [2] picture header bits (MSB first), pic_hdr_bits long -- in device memory, so that a
                               // captured CUDA graph of the frame step can be replayed with another temporal reference;
  uint32_t pic_hdr_imm[2];     // NULL: the bits are these (fixed quantiser: nothing is captured)
  int pic_hdr_bits;
  int gquant;                  // reported when rc.rate == 0
  RcArgs rc;
};
__device__ __forceinline__ uint32_t* vlc_out_u32(uint8_t* out, int S, int field) { return reinterpret_cast<uint32_t*>(out) + (field == 0 ? 0 : (S + 1) + (field - 1) * S); }
__device__ __forceinline__ unsigned long long* vlc_out_bitpos(uint8_t* out, int S) {
  return reinterpret_cast<unsigned long long*>(out + vlc_bitpos_offset(S));
}
__device__ __forceinline__ uint32_t* vlc_out_rc(uint8_t* out, int S, int field) {     // 0: gquant, 1: overflows
  return reinterpret_cast<uint32_t*>(out + vlc_bitpos_offset(S) + (size_t)S * 8) + field * S;
}

// chunk sizes and their packed offsets: one CTA; four threads sum a stream's piece lengths (independent loads), then a scan
// over the streams, in blocks of blockDim.x / 4 streams with a running base
__global__ void __launch_bounds__(1024)
vlc_sizes_kernel(const __grid_constant__ VlcFrameArgs a) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* off = vlc_out_u32(a.out, a.n_streams, 0);
  uint32_t* nb = vlc_out_u32(a.out, a.n_streams, 1);
  const int np = a.ngob * a.ppg, per = blockDim.x >> 2;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int s0 = 0; s0 < a.n_streams; s0 += per) {
    const int s = s0 + (tid >> 2), q = tid & 3;
    uint32_t bits = 0;
    if (s < a.n_streams) {
      const uint32_t* gb = a.gob_bits + (size_t)s * np;
#pragma unroll
      for (int i = 0; i < (12 * VLC_PPG + 3) / 4; i++) if (q + 4 * i < np) bits += gb[q + 4 * i];     // (all in flight at once)
    }
    bits += __shfl_xor_sync(0xffffffffu, bits, 1);
    bits += __shfl_xor_sync(0xffffffffu, bits, 2);
    uint32_t x = 0;
    if (s < a.n_streams && q == 0) {
      const uint32_t bytes = (bits + a.carry_len[s] + (uint32_t)a.pic_hdr_bits) >> 3;
      nb[s] = bytes;
      x = (bytes + 3u) & ~3u;
    }
    const uint32_t mine = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint32_t base = s_base;
    for (int w = 0; w < warp; w++) base += s_w[w];
    if (s < a.n_streams && q == 0) off[s] = base + x - mine;
    __syncthreads();
    if (tid == blockDim.x - 1) s_base = base + x;
    __syncthreads();
  }
  if (tid == 0) off[a.n_streams] = s_base;
}

// One CTA per stream: the carry bits of the previous frame, the picture header and the pieces are concatenated into the
// frame's byte chunk.  An output word that lies inside one piece is a funnel shift of two adjacent words of that piece
// with a shift that is the same for the whole piece: the warps take pieces in turn and stream them (coalesced loads and
// stores, no search).  The words that hold a piece boundary -- at most one per piece -- are assembled bit by bit.
constexpr int VLC_FRAME_THREADS = 1024;              // 32 warps: a warp's pieces are a chain of dependent round trips to L2
__global__ void __launch_bounds__(VLC_FRAME_THREADS)
vlc_frame_kernel(const __grid_constant__ VlcFrameArgs a) {
  constexpr int MAXP = 2 + 12 * VLC_PPG;
  static_assert(MAXP <= 128, "the offsets are scanned by four warps");
  __shared__ uint32_t s_off[MAXP + 1];
  __shared__ uint32_t s_small[2][3];                  // piece 0 (carry) and 1 (picture header) as word arrays
  __shared__ uint32_t s_wsum[4];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ng = a.ngob * a.ppg, np = 2 + ng;
  {
    uint32_t len = 0;
    if (tid == 0) len = a.carry_len[s];
    else if (tid == 1) len = (uint32_t)a.pic_hdr_bits;
    else if (tid < np) len = a.gob_bits[(size_t)s * ng + (tid - 2)];
    uint32_t x = len;
    if (warp < 4) {
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
      if (lane == 31) s_wsum[warp] = x;
    }
    if (tid == 0) {
      s_small[0][0] = a.carry[s]; s_small[0][1] = 0; s_small[0][2] = 0;
      s_small[1][0] = a.pic_hdr ? a.pic_hdr[0] : a.pic_hdr_imm[0]; s_small[1][1] = a.pic_hdr ? a.pic_hdr[1] : a.pic_hdr_imm[1]; s_small[1][2] = 0;
    }
    __syncthreads();
    if (warp < 4) {
      uint32_t base = 0;
      for (int w = 0; w < warp; w++) base += s_wsum[w];
      if (tid < np) s_off[tid] = base + x - len;
      if (tid == 127) s_off[np] = base + x;
    }
  }
  __syncthreads();
  const uint32_t total = s_off[np];
  const uint32_t* slot0 = a.gob_words + (size_t)s * a.ngob * VLC_GOB_WORDS;
  auto piece_words = [&](int p) -> const uint32_t* {
    const int q = p - 2;
    return p < 2 ? s_small[p] : slot0 + (size_t)(q / a.ppg) * VLC_GOB_WORDS + (a.ppg > 1 ? vlc_piece_m0(q % a.ppg) * VLC_MB_WORDS : 0);
  };
  // 32 bits of the concatenation starting at bit o (zeros beyond the end)
  auto get32 = [&](uint32_t o) -> uint32_t {
    int p = 0, hi = np;                                // the last piece that starts at or before o
    while (hi - p > 1) { const int mid = (p + hi) >> 1; if (s_off[mid] <= o) p = mid; else hi = mid; }
    uint32_t res = 0;
    int filled = 0;
    while (filled < 32 && p < np) {
      const int avail = (int)(s_off[p + 1] - o);
      if (avail <= 0) { p++; continue; }
      const int take = min(32 - filled, avail);
      const uint32_t local = o - s_off[p];
      const uint32_t* src = piece_words(p);
      uint32_t v = __funnelshift_l(src[(local >> 5) + 1], src[local >> 5], local & 31);
      if (take < 32) v &= ~(0xffffffffu >> take);
      res |= v >> filled;
      filled += take; o += (uint32_t)take;
    }
    return res;
  };
  const uint32_t nbytes = total >> 3, nwords = (nbytes + 3) >> 2;
  uint32_t* out = reinterpret_cast<uint32_t*>(a.out + vlc_data_offset(a.n_streams) + vlc_out_u32(a.out, a.n_streams, 0)[s]);
  // (stream order = big endian)
  for (int p = warp; p < np; p += VLC_FRAME_THREADS / 32) {    // words inside one piece
    const uint32_t o = s_off[p], e = s_off[p + 1];
    const uint32_t w0 = (o + 31) >> 5, w1 = min(e >> 5, nwords);
    if (w1 <= w0) continue;
    const uint32_t* src = piece_words(p);                      // output word w0 + i starts in word i of the piece, at bit `sh`
    const uint32_t sh = (0u - o) & 31, n = w1 - w0;
#pragma unroll 4
    for (uint32_t i = lane; i < n; i += 32) out[w0 + i] = __byte_perm(__funnelshift_l(src[i + 1], src[i], sh), 0, 0x0123);
  }
  if (tid < np) {                                              // the word in which piece `tid` ends, if it ends inside a word
    const uint32_t o = s_off[tid], e = s_off[tid + 1];
    if (e > o && (e & 31) && (e >> 5) < nwords) out[e >> 5] = __byte_perm(get32(e & ~31u), 0, 0x0123);
  }
  if (tid == 0) {
    const uint32_t rem = total & 7;
    const uint32_t cbits = rem ? (get32(8 * nbytes) & ~(0xffffffffu >> rem)) : 0u;
    const unsigned long long pos = a.bitpos[s] + (total - s_off[1]);      // the carried-in bits were counted with their own frame
    a.carry[s] = cbits; a.carry_len[s] = rem; a.bitpos[s] = pos;
    vlc_out_u32(a.out, a.n_streams, 2)[s] = cbits;
    vlc_out_u32(a.out, a.n_streams, 3)[s] = rem;
    vlc_out_bitpos(a.out, a.n_streams)[s] = pos;
    uint32_t gq = (uint32_t)a.gquant, novf = 0;
    if (a.rc.rate) {                                   // end of frame, p64.c:670-680
      long long bo = a.rc.buffer_offset[s];
      if (a.rc.first_frame) bo = (a.rc.rate / 4) / 2 - rc_buffer_contents(a.rc, (long long)pos, bo, a.ngob, 0);
      bo -= (int)((long long)a.rc.rate * a.rc.frame_skip * a.rc.frame_rate_div) / a.rc.frame_rate;     // the product wraps in C int BEFORE the division (p64.c:677)
      a.rc.buffer_offset[s] = bo;
      gq = a.rc.quant[s]; novf = a.rc.overflows[s];
    }
    vlc_out_rc(a.out, a.n_streams, 0)[s] = gq;
    vlc_out_rc(a.out, a.n_streams, 1)[s] = novf;
  }
}

}  // namespace p64b
