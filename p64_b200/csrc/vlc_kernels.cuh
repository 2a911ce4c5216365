// Device-side entropy coding: H.261 headers and run-level VLC on the GPU, so that a fixed-quantiser
// frame step returns the finished bit stream instead of records + levels.  Bit-exact restatement of the host coder
// (bits.cpp), i.e. of the reference's
//   WriteGOBHeader  marker.c:182-209      WriteMBHeader  marker.c:288-354 (MVD prediction 310-338)
//   EncodeDC / EncodeAC / CBPEncodeAC  codec.c:96-205, 346-355      WritePictureHeader  marker.c:103-137
//   mputv  stream.c:193-205 (MSB first)
//
//   vlc_gob_seq_kernel  fixed quantiser: one WARP per (stream, GOB), macroblock after macroblock, one item (intra DC, level,
//                     EOB) per lane, the words streamed out through a small ring in shared memory -- no measuring pass.
//   vlc_gob_kernel    rate control: one CTA per (stream, GOB): one thread per piece (GOB header, 33 x {MB header, 6 blocks});
//                     pass 1 measures every piece, a CTA-wide scan places them, pass 2 writes the bits into a
//                     shared-memory image of the GOB, which is then stored unshifted into the GOB's scratch slot.
//   vlc_frame_kernel  one CTA per stream: carry bits of the previous frame + picture header + the GOB strings are
//                     gathered word by word (funnel shifts) into the frame's byte chunk; the < 8 trailing bits stay
//                     on the device as the next frame's carry, so the host only ever appends whole bytes.
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
constexpr int VLC_GOB_WORDS = (26 + 33 * (VLC_MBHDR_MAX_BITS + 6 * VLC_BLOCK_MAX_BITS) + 31) / 32 + 8;   // slot size, words
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
  uint32_t* gob_words;       // [S][ngob][VLC_GOB_WORDS] scratch
  uint32_t* gob_bits;        // [S][ngob]
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
// Fixed quantiser (round 2): ONE WARP PER GOB, macroblock after macroblock, one ITEM per lane.
// A macroblock's bit string is a header followed by items in a fixed order: per coded block [intra DC], the non-zero levels in
// transmission order, [EOB].  The warp that owns a GOB knows the bit position it has reached, so nothing has to be measured
// first and placed later: per macroblock
//   list   lanes 1..24 (block b, quarter q) stage their 16 levels in shared memory and write the descriptors of their items
//          (block | kind | position) into the warp's list at positions given by a warp scan of the counts: stream order;
//   code   lane k takes item k, 32 at a time: the run comes from the previous descriptor, the code from the table in shared
//          memory (codec.c:96-205: escapes as in 113-115, the two-bit code for a leading +-1 of a CBP-coded block; EncodeDC
//          codec.c:346-355); a warp scan of the lengths places the items behind the header;
//   emit   every lane ORs its item into a 64-word ring in shared memory; the words the position has passed are stored to the
//          GOB's scratch slot and zeroed.
// The next macroblock's record and levels are loaded while the current one is coded.  No CTA-wide barrier after the table
// load, no image of the GOB in shared memory (1.9 KB per warp instead of 34 KB per CTA), every level is looked at once.
// Measured on 256 CIF frames: 57.9 M warp-instructions, 0.108 ms (profiles/r02_ncu_vlc_gob_seq_kernel.txt).
// The thread-per-piece kernel below (round 1) measures every piece, scans, and walks the levels again to write them, waiting
// for the busiest block of every warp in both walks and at five CTA barriers: 71 M warp-instructions per 256 CIF frames,
// 0.116 ms.  It remains the rate-control kernel (one GOB per stream per launch: 256 GOBs cannot feed 148 SMs with one warp
// each).  Two CTA-per-GOB re-mappings tried on the way were no faster than it (a lane per block quarter: the zig-zag order
// puts the non-zero levels into the first quarter, 121 M / 0.149 ms; an item per lane with the codes kept in registers between
// a measuring and an emitting phase: 70 M / 0.116 ms, 3.9 barrier stalls per issue) -- profiles/r02_ncu_vlc_gob_kernel_*.
constexpr int VLC_SEQ_WARPS = 4;
constexpr int VLC_SEQ_THREADS = 32 * VLC_SEQ_WARPS;
constexpr int VLC_ITEMS_MAX = 6 * 66;                   // per macroblock: 6 x (DC or position 0, 63 levels, EOB) at most
constexpr int VLC_RING = 64;                            // words; one sub-round adds at most 47 + 32 x 20 bits = 22 words

// list: stages the macroblock's levels and writes its item descriptors in stream order; returns the number of items
__device__ __forceinline__ int vlc_list_items(const p64b_mb& rec, uint4 v, int lane, uint8_t* stage, uint16_t* items) {
  const int mt = rec.mtype;
  const bool cbp_type = vt(V_CBP, mt);
  const int b = (lane - 1) >> 2, q = (lane - 1) & 3;
  const bool blk = lane >= 1 && lane <= 24;
  const bool coded = blk && vt(V_TCOEF, mt) && ((rec.cbp >> (5 - b)) & 1);
  if (!coded) v = make_uint4(0, 0, 0, 0);               // (the levels were loaded before the record was known)
  if (blk) *reinterpret_cast<uint4*>(stage + (lane - 1) * 16) = v;
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t nz = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t hi = (w[j] | ((w[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;     // bit 7 of every non-zero byte
    nz |= ((((hi >> 7) * 0x01020408u) >> 24) & 0xfu) << (4 * j);
  }
  if (q == 0 && !cbp_type) nz &= ~1u;                      // the intra DC is its own item
  const uint32_t anym = __ballot_sync(0xffffffffu, coded && nz != 0);
  const bool dc = coded && q == 0 && !cbp_type;
  const bool eob = coded && q == 3 && (!cbp_type || ((anym >> (1 + 4 * b)) & 0xfu));       // an all-zero CBP block gets no EOB (codec.c:169-174)
  const int cnt = __popc(nz) + (dc ? 1 : 0) + (eob ? 1 : 0);
  int x = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  const int total = __shfl_sync(0xffffffffu, x, 31);
  int idx = x - cnt;
  const uint32_t base = (uint32_t)b << 8;
  if (dc) items[idx++] = (uint16_t)(base | (1u << 6));
  while (nz) {
    const int pl = __ffs((int)nz) - 1;
    nz &= nz - 1;
    items[idx++] = (uint16_t)(base | (uint32_t)(16 * q + pl));
  }
  if (eob) items[idx++] = (uint16_t)(base | (2u << 6) | 63u);
  __syncwarp();
  return total;
}

// code: item k of the list -> (code bits, length)
__device__ __forceinline__ void vlc_code_item(int k, int total, const uint16_t* items, const uint8_t* stage, const uint32_t* s_tcoef,
                                              uint32_t& code, int& len) {
  code = 0; len = 0;
  if (k >= total) return;
  const uint32_t d = items[k];
  const int b = (int)(d >> 8), kind = (int)((d >> 6) & 3u), p = (int)(d & 63u);
  if (kind == 1) {                                     // EncodeDC, codec.c:346-355
    int dc = stage[b * 64];
    dc = min(max(dc, 1), 254);
    if (dc == 128) dc = 255;
    code = (uint32_t)dc; len = 8;
  } else if (kind == 2) {                              // EOB "10"
    code = 2u; len = 2;
  } else {
    const int v = (int)(int8_t)stage[b * 64 + p], a = abs(v);
    int prevp = -1;                                    // the previous item of the same block is its DC (position 0) or a level
    if (k > 0) { const uint32_t pd = items[k - 1]; if ((int)(pd >> 8) == b) prevp = (int)(pd & 63u); }
    const int run = p - prevp - 1;
    if (p == 0 && a == 1) {                            // "1s": a leading +-1 of a CBP-coded block (only those list position 0)
      code = 2u | (uint32_t)(v < 0); len = 2;
    } else {
      const uint32_t e = (run < 32 && a < 16) ? s_tcoef[run * 16 + a] : 0u;
      if (e) { len = (int)(e >> 16) + 1; code = ((e & 0xffffu) << 1) | (uint32_t)(v < 0); }
      else { code = (1u << 14) | ((uint32_t)run << 8) | (uint32_t)(v & 0xff); len = 20; }     // escape (codec.c:113-115)
    }
  }
}


// OR `n` bits (n <= 32, v < 2^n) into the ring of VLC_RING zeroed words that holds the MSB-first GOB string around the ABSOLUTE
// bit offset `o`
__device__ __forceinline__ void vlc_or_ring(uint32_t* ring, uint32_t o, uint32_t v, int n) {
  if (!n) return;
  const uint32_t w = o >> 5;
  const int sh = (int)(o & 31), over = sh + n - 32;
  if (over <= 0) atomicOr(ring + (w & (VLC_RING - 1)), v << (-over));
  else { atomicOr(ring + (w & (VLC_RING - 1)), v >> over); atomicOr(ring + ((w + 1) & (VLC_RING - 1)), v << (32 - over)); }
}

__global__ void __launch_bounds__(VLC_SEQ_THREADS)
vlc_gob_seq_kernel(const __grid_constant__ VlcArgs a) {
  __shared__ DevVlcTables s_t;
  __shared__ __align__(16) uint8_t s_stage[VLC_SEQ_WARPS][384];
  __shared__ uint16_t s_items[VLC_SEQ_WARPS][VLC_ITEMS_MAX + 4];
  __shared__ uint32_t s_ring[VLC_SEQ_WARPS][VLC_RING];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < DEV_VLC_WORDS; i += VLC_SEQ_THREADS) reinterpret_cast<uint32_t*>(&s_t)[i] = reinterpret_cast<const uint32_t*>(a.tables)[i];
  for (int i = lane; i < VLC_RING; i += 32) s_ring[warp][i] = 0;
  __syncthreads();
  const int task = blockIdx.x * VLC_SEQ_WARPS + warp;
  if (task >= a.n_streams * a.gob_count) return;        // (no barrier below)
  const int s = task / a.gob_count, gob = a.gob_first + task % a.gob_count;
  const size_t mb0 = (size_t)s * a.nmb + gob * 33;
  uint8_t* stage = s_stage[warp];
  uint16_t* items = s_items[warp];
  uint32_t* ring = s_ring[warp];
  uint32_t* dst = a.gob_words + ((size_t)s * a.ngob + gob) * VLC_GOB_WORDS;
  const int8_t* lv0 = a.levels + mb0 * P64B_LEVELS_PER_MB;
  const bool blk = lane >= 1 && lane <= 24;

  uint32_t pos = 26, wbase = 0;                        // bits written so far; first word of the GOB string still in the ring
  if (lane == 0) {                                     // WriteGOBHeader, marker.c:182-209: GBSC, GN, GQUANT, no GSPARE
    const int gn = (a.qcif ? (gob << 1) : gob) + 1;
    vlc_or_ring(ring, 0, (1u << 10) | ((uint32_t)gn << 6) | ((uint32_t)a.gquant << 1), 26);
  }
  // the words [wbase, pos / 32) are final: store them and hand their ring slots back
  auto flush = [&]() {
    __syncwarp();
    const uint32_t n = (pos >> 5) - wbase;              // <= 23 < 32
    if ((uint32_t)lane < n) { const uint32_t w = wbase + (uint32_t)lane; dst[w] = ring[w & (VLC_RING - 1)]; ring[w & (VLC_RING - 1)] = 0; }
    wbase += n;
    __syncwarp();
  };
  p64b_mb prev{};
  uint2 rec_next = __ldg(reinterpret_cast<const uint2*>(a.mbs + mb0));
  uint4 lv_next = make_uint4(0, 0, 0, 0);
  if (blk) lv_next = __ldg(reinterpret_cast<const uint4*>(lv0 + (lane - 1) * 16));
#pragma unroll 1
  for (int m = 0; m < 33; m++) {
    const uint2 rw = rec_next;
    const uint4 lv = lv_next;
    if (m + 1 < 33) {                                   // the next macroblock's record and levels: in flight while this one is coded
      rec_next = __ldg(reinterpret_cast<const uint2*>(a.mbs + mb0 + m + 1));
      if (blk) lv_next = __ldg(reinterpret_cast<const uint4*>(lv0 + (size_t)(m + 1) * P64B_LEVELS_PER_MB + (lane - 1) * 16));
    }
    const p64b_mb rec = *reinterpret_cast<const p64b_mb*>(&rw);
    int hl;
    const uint64_t hb = vlc_mb_header(rec, prev, m, &s_t, &hl);
    prev = rec;
    if (lane == 0) {
      if (hl > 32) { vlc_or_ring(ring, pos, (uint32_t)(hb >> 32), hl - 32); vlc_or_ring(ring, pos + (uint32_t)(hl - 32), (uint32_t)hb, 32); }
      else vlc_or_ring(ring, pos, (uint32_t)hb, hl);
    }
    pos += (uint32_t)hl;
    const int total = vlc_list_items(rec, lv, lane, stage, items);
    if (total == 0) flush();
#pragma unroll 1
    for (int k0 = 0; k0 < total; k0 += 32) {
      uint32_t code; int len;
      vlc_code_item(k0 + lane, total, items, stage, s_t.tcoef, code, len);
      int x = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
      vlc_or_ring(ring, pos + (uint32_t)(x - len), code, len);
      pos += (uint32_t)__shfl_sync(0xffffffffu, x, 31);
      flush();
    }
  }
  __syncwarp();
  const uint32_t nwords = (pos + 31) >> 5;               // the partial last word, + one zero word for the gather's look-ahead
  if (lane == 0) {
    if (pos & 31) dst[pos >> 5] = ring[(pos >> 5) & (VLC_RING - 1)];
    dst[nwords] = 0;
    a.gob_bits[(size_t)s * a.ngob + gob] = pos;
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
  const uint32_t* pic_hdr;     // [2] picture header bits (MSB first), pic_hdr_bits long -- in device memory, so that a
                               // captured CUDA graph of the frame step can be replayed with another temporal reference
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

// chunk sizes and their packed offsets: one CTA, streams in blocks of blockDim.x with a running base
__global__ void __launch_bounds__(1024)
vlc_sizes_kernel(const __grid_constant__ VlcFrameArgs a) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* off = vlc_out_u32(a.out, a.n_streams, 0);
  uint32_t* nb = vlc_out_u32(a.out, a.n_streams, 1);
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int s0 = 0; s0 < a.n_streams; s0 += blockDim.x) {
    const int s = s0 + tid;
    uint32_t bytes = 0;
    if (s < a.n_streams) {
      uint32_t bits = a.carry_len[s] + (uint32_t)a.pic_hdr_bits;
      for (int g = 0; g < a.ngob; g++) bits += a.gob_bits[(size_t)s * a.ngob + g];
      bytes = bits >> 3;
      nb[s] = bytes;
    }
    uint32_t x = (bytes + 3u) & ~3u;
    const uint32_t mine = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint32_t base = s_base;
    for (int w = 0; w < warp; w++) base += s_w[w];
    if (s < a.n_streams) off[s] = base + x - mine;
    __syncthreads();
    if (tid == blockDim.x - 1) s_base = base + x;
    __syncthreads();
  }
  if (tid == 0) off[a.n_streams] = s_base;
}

__global__ void __launch_bounds__(VLC_THREADS)
vlc_frame_kernel(const __grid_constant__ VlcFrameArgs a) {
  constexpr int MAXP = 2 + 12;
  __shared__ uint32_t s_off[MAXP + 1];
  __shared__ uint32_t s_small[2][3];                  // piece 0 (carry) and 1 (picture header) as word arrays
  const int s = blockIdx.x, tid = threadIdx.x, np = 2 + a.ngob;
  if (tid == 0) {
    uint32_t o = 0;
    s_off[0] = 0; o += a.carry_len[s];
    s_off[1] = o; o += (uint32_t)a.pic_hdr_bits;
    for (int g = 0; g < a.ngob; g++) { s_off[2 + g] = o; o += a.gob_bits[(size_t)s * a.ngob + g]; }
    s_off[np] = o;
    s_small[0][0] = a.carry[s]; s_small[0][1] = 0; s_small[0][2] = 0;
    s_small[1][0] = a.pic_hdr[0]; s_small[1][1] = a.pic_hdr[1]; s_small[1][2] = 0;
  }
  __syncthreads();
  const uint32_t total = s_off[np];
  // 32 bits of the concatenation starting at bit o (zeros beyond the end)
  auto get32 = [&](uint32_t o) -> uint32_t {
    int p = 0;
    while (p < np - 1 && s_off[p + 1] <= o) p++;
    uint32_t res = 0;
    int filled = 0;
    while (filled < 32 && p < np) {
      const int avail = (int)(s_off[p + 1] - o);
      if (avail <= 0) { p++; continue; }
      const int take = min(32 - filled, avail);
      const uint32_t local = o - s_off[p];
      const uint32_t* src = p < 2 ? s_small[p] : a.gob_words + ((size_t)s * a.ngob + (p - 2)) * VLC_GOB_WORDS;
      uint32_t v = __funnelshift_l(src[(local >> 5) + 1], src[local >> 5], local & 31);
      if (take < 32) v &= ~(0xffffffffu >> take);
      res |= v >> filled;
      filled += take; o += (uint32_t)take;
    }
    return res;
  };
  const uint32_t nbytes = total >> 3, nwords = (nbytes + 3) >> 2;
  uint32_t* out = reinterpret_cast<uint32_t*>(a.out + vlc_data_offset(a.n_streams) + vlc_out_u32(a.out, a.n_streams, 0)[s]);
  for (uint32_t w = tid; w < nwords; w += VLC_THREADS) out[w] = __byte_perm(get32(32 * w), 0, 0x0123);   // stream order = big endian
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
