// Device-side entropy coding: H.261 headers and run-level VLC on the GPU, so that a fixed-quantiser
// frame step returns the finished bit stream instead of records + levels.  Bit-exact restatement of the host coder
// (bits.cpp), i.e. of the reference's
//   WriteGOBHeader  marker.c:182-209      WriteMBHeader  marker.c:288-354 (MVD prediction 310-338)
//   EncodeDC / EncodeAC / CBPEncodeAC  codec.c:96-205, 346-355      WritePictureHeader  marker.c:103-137
//   mputv  stream.c:193-205 (MSB first)
//
//   vlc_gob_kernel    one CTA per (stream, GOB): one thread per piece (GOB header, 33 x {MB header, 6 blocks});
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

struct VlcArgs {
  const DevVlcTables* tables;
  const p64b_mb* mbs;        // [S][nmb] GOB-major
  const int8_t* levels;      // [S][nmb][6][64] transmission order
  uint32_t* gob_words;       // [S][ngob][VLC_GOB_WORDS] scratch
  uint32_t* gob_bits;        // [S][ngob]
  int n_streams, ngob, nmb;
  int qcif;                  // GOB numbers 1,3,5 (p64.c:709-711)
  int gquant;
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
template <bool EMIT>
__device__ __forceinline__ int vlc_block(const int8_t* __restrict__ lv, bool cbp_type, const uint32_t* s_tcoef, BitEmitter* em) {
  // non-zero mask of the 64 levels
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

__global__ void __launch_bounds__(VLC_THREADS)
vlc_gob_kernel(const __grid_constant__ VlcArgs a) {
  __shared__ uint32_t s_buf[VLC_GOB_WORDS];
  __shared__ DevVlcTables s_t;
  __shared__ uint32_t s_off[VLC_THREADS + 1];
  __shared__ uint32_t s_wsum[VLC_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x / a.ngob, gob = blockIdx.x % a.ngob;
  for (int i = tid; i < DEV_VLC_WORDS; i += VLC_THREADS) reinterpret_cast<uint32_t*>(&s_t)[i] = reinterpret_cast<const uint32_t*>(a.tables)[i];
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
  const int8_t* lv = a.levels + (mbi * 6 + (k ? k - 1 : 0)) * 64;
  if (is_piece) {
    if (piece == 0) {                                 // WriteGOBHeader, marker.c:182-209: GBSC, GN, GQUANT, no GSPARE
      const int gn = (a.qcif ? (gob << 1) : gob) + 1;
      hbits = (1ull << 10) | ((uint64_t)gn << 6) | ((uint64_t)a.gquant << 1);
      len = 26;
    } else {
      rec = a.mbs[mbi];
      if (k == 0) {
        if (m) prev = a.mbs[mbi - 1];
        hbits = vlc_mb_header(rec, prev, m, &s_t, &len);
      } else {
        coded = vt(V_TCOEF, rec.mtype) && ((rec.cbp >> (6 - k)) & 1);      // block c = k-1: bit 5-c
        if (coded) len = vlc_block<false>(lv, vt(V_CBP, rec.mtype), s_t.tcoef, nullptr);
      }
    }
  }

  // ---- exclusive scan of the piece lengths
  uint32_t x = (uint32_t)len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) s_wsum[warp] = x;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; w++) base += s_wsum[w];
  const uint32_t off = base + x - (uint32_t)len;
  if (tid == VLC_THREADS - 1) s_off[0] = base + x;     // total bits of the GOB
  __syncthreads();
  const uint32_t total = s_off[0], nwords = (total + 31) >> 5;
  for (uint32_t i = tid; i < nwords + 1; i += VLC_THREADS) s_buf[i] = 0;
  __syncthreads();

  // ---- pass 2: the bits
  if (len) {
    BitEmitter em(s_buf, off);
    if (k == 0) {
      if (len > 32) { em.put((uint32_t)(hbits >> 32), len - 32); em.put((uint32_t)hbits, 32); }
      else em.put((uint32_t)hbits, len);
    } else {
      vlc_block<true>(lv, vt(V_CBP, rec.mtype), s_t.tcoef, &em);
    }
    em.flush();
  }
  __syncthreads();
  uint32_t* dst = a.gob_words + ((size_t)s * a.ngob + gob) * VLC_GOB_WORDS;
  for (uint32_t i = tid; i < nwords + 1; i += VLC_THREADS) dst[i] = s_buf[i];      // + one zero word for the gather's look-ahead
  if (tid == 0) a.gob_bits[(size_t)s * a.ngob + gob] = total;
}

// Output of one frame step, one buffer (copied to the host in one piece):
//   uint32 offset[S+1]   byte offset of stream s's chunk in data[] (multiples of 4); offset[S] = bytes used
//   uint32 nbytes[S]     whole bytes of the chunk
//   uint32 carry[S]      the stream's pending bits (< 8) after this frame, left-aligned in the word
//   uint32 carry_len[S]
//   uint64 bitpos[S]     bits written so far (mwtell, stream.c:233-238)
//   uint8  data[]        at byte vlc_data_offset(S)
__host__ __device__ constexpr size_t vlc_data_offset(int S) { return (((size_t)(4 * S + 1) * 4 + 15) / 16) * 16 + (size_t)S * 8; }
struct VlcFrameArgs {
  const uint32_t* gob_words;   // [S][ngob][VLC_GOB_WORDS]
  const uint32_t* gob_bits;    // [S][ngob]
  uint32_t* carry;             // [S] stream state, resident
  uint32_t* carry_len;         // [S]
  unsigned long long* bitpos;  // [S]
  uint8_t* out;                // the step's output buffer (layout above)
  int n_streams, ngob;
  uint32_t pic_hdr[2];         // picture header bits (MSB first), pic_hdr_bits long
  int pic_hdr_bits;
};
__device__ __forceinline__ uint32_t* vlc_out_u32(uint8_t* out, int S, int field) { return reinterpret_cast<uint32_t*>(out) + (field == 0 ? 0 : (S + 1) + (field - 1) * S); }
__device__ __forceinline__ unsigned long long* vlc_out_bitpos(uint8_t* out, int S) {
  return reinterpret_cast<unsigned long long*>(out + (((size_t)(4 * S + 1) * 4 + 15) / 16) * 16);
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
  }
}

}  // namespace p64b
