// Ingest: what the reference's Y4M reader does to a frame before the encoder sees it
// (y4m_input.c:195-545, selected in y4m_input_open_impl y4m_input.c:587-655), as one gather kernel over the uploaded chroma
// payload.  The luma plane never needs conversion (it is uploaded straight into the source frame); the chroma planes
// are re-sited / re-sampled to 4:2:0 "jpeg" siting:
//   420jpeg / 420   none (y4m_convert_null)
//   420mpeg2        horizontal quarter-pel shift, 6-tap [4 -17 114 35 -9 1]/128 (y4m_input.c:195-229)
//   420paldv        the same horizontal filter into an 8-bit intermediate, then a vertical quarter-pel shift:
//                   Cb up [1 -9 35 114 -17 4]/128, Cr down [4 -17 114 35 -9 1]/128 (y4m_input.c:274-376)
//   422             the 420mpeg2 filter on full-height planes (y4m_input.c:617-624 selects the same function)
//   411             4:1:1 -> 4:2:2 by [1 110 18 -1]/128 and [-3 50 86 -5]/128 (y4m_input.c:417-459)
//   444 / 444alpha  none; mono: chroma = 128 (y4m_input.c:463-471)
// All sums are integer, rounded (s+64)>>7 with an arithmetic shift and clamped to [0,255]; every tap index is clamped
// to the plane edge (the reference's three loops per row are exactly that).
// What the ENCODER then reads (ReadIob io.c:636-645 installs plane pointers, ReadBlock io.c:793-803 walks them with the
// CIF/QCIF stride Iob->width): the first (W/2)*(H/2) bytes of each converted chroma plane, whatever its real shape.  For
// 4:2:2 / 4:1:1 that is the top half of the plane, for 4:4:4 the first quarter of its bytes -- reproduced as is
// ("bug-compatible"), because the criterion is the byte-identical stream.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/p64_b200.h"

namespace p64b {

struct IngestArgs {
  const uint8_t* aux;   // [S][aux_stride] uploaded chroma payload (Cb plane, Cr plane)
  size_t aux_stride;
  uint8_t* dst;         // [S][dst_stride] source frames; chroma written at W*H
  size_t dst_stride;
  int W, H, n_streams, chroma;
};

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int byte_of(uint32_t w, int k) { return (int)((w >> (8 * k)) & 0xffu); }
__device__ __forceinline__ uint32_t pack_u8x4(const int (&o)[4]) {
  return (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
}

// Samples x0-4 .. x0+7 of a row of n bytes (n and x0 multiples of 4, row 4-byte aligned) as 12 ints, indices clamped to
// the row: three aligned word loads; beyond the ends the edge sample is replicated.
template <typename Load>
__device__ __forceinline__ void row12(Load ld, int x0, int n, int (&p)[12]) {
  const uint32_t w1 = ld(x0 >> 2);
  const uint32_t w0 = x0 ? ld((x0 >> 2) - 1) : (w1 & 0xffu) * 0x01010101u;
  const uint32_t w2 = x0 + 4 < n ? ld((x0 >> 2) + 1) : (w1 >> 24) * 0x01010101u;
#pragma unroll
  for (int k = 0; k < 4; k++) { p[k] = byte_of(w0, k); p[4 + k] = byte_of(w1, k); p[8 + k] = byte_of(w2, k); }
}
// [4 -17 114 35 -9 1]/128 at x0 .. x0+3 (first tap at x-2): the quarter-pel shift of y4m_input.c:209-224
__device__ __forceinline__ void hshift6x4(const int (&p)[12], int (&o)[4]) {
#pragma unroll
  for (int j = 0; j < 4; j++)
    o[j] = clamp255((4 * p[j + 2] - 17 * p[j + 3] + 114 * p[j + 4] + 35 * p[j + 5] - 9 * p[j + 6] + p[j + 7] + 64) >> 7);
}

// One CTA per (stream, chroma plane).  4 output samples per thread and step; word loads only.  420paldv keeps the whole
// horizontally filtered plane (8-bit, as the reference's intermediate buffer) in shared memory -- 25 KB for CIF -- and
// filters it vertically from there, so every intermediate sample is computed once.
constexpr int INGEST_THREADS = 256;
__global__ void __launch_bounds__(INGEST_THREADS) ingest_chroma_kernel(const IngestArgs a) {
  extern __shared__ __align__(16) uint32_t s_tmp[];                      // paldv only: [ch][cw] bytes
  const int cw = a.W >> 1, ch = a.H >> 1, q = cw >> 2, nwords = ch * q;
  const int s = blockIdx.x >> 1, pl = blockIdx.x & 1;
  const uint8_t* aux = a.aux + (size_t)s * a.aux_stride;
  uint32_t* dst = reinterpret_cast<uint32_t*>(a.dst + (size_t)s * a.dst_stride + (size_t)a.W * a.H + (size_t)pl * cw * ch);
  switch (a.chroma) {
    case P64B_CHROMA_420MPEG2:
    case P64B_CHROMA_422:
    case P64B_CHROMA_420PALDV: {
      const int sh = a.chroma == P64B_CHROMA_422 ? a.H : ch;            // source plane height; only its first ch rows are read
      const uint32_t* plane = reinterpret_cast<const uint32_t*>(aux + (size_t)pl * cw * sh);
      uint32_t* hout = a.chroma == P64B_CHROMA_420PALDV ? s_tmp : dst;
      for (int i = threadIdx.x; i < nwords; i += INGEST_THREADS) {
        const int y = i / q, x0 = (i - y * q) * 4;
        const uint32_t* row = plane + y * q;
        int p[12], o[4];
        row12([&](int w) { return __ldg(row + w); }, x0, cw, p);
        hshift6x4(p, o);
        hout[i] = pack_u8x4(o);
      }
      if (a.chroma != P64B_CHROMA_420PALDV) break;
      __syncthreads();
      // vertical quarter-pel shift of the intermediate: Cb up [1 -9 35 114 -17 4] (first tap at y-3), Cr down
      // [4 -17 114 35 -9 1] (first tap at y-2); row indices clamped (y4m_input.c:313-359)
      for (int i = threadIdx.x; i < nwords; i += INGEST_THREADS) {
        const int y = i / q, xw = i - y * q;
        uint32_t t[6];
        const int first = pl == 0 ? y - 3 : y - 2;
#pragma unroll
        for (int k = 0; k < 6; k++) t[k] = s_tmp[min(max(first + k, 0), ch - 1) * q + xw];
        int o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int v0 = byte_of(t[0], j), v1 = byte_of(t[1], j), v2 = byte_of(t[2], j), v3 = byte_of(t[3], j), v4 = byte_of(t[4], j), v5 = byte_of(t[5], j);
          o[j] = pl == 0 ? clamp255((v0 - 9 * v1 + 35 * v2 + 114 * v3 - 17 * v4 + 4 * v5 + 64) >> 7)
                         : clamp255((4 * v0 - 17 * v1 + 114 * v2 + 35 * v3 - 9 * v4 + v5 + 64) >> 7);
        }
        dst[i] = pack_u8x4(o);
      }
      break;
    }
    case P64B_CHROMA_411: {
      const int sw = (a.W + 3) >> 2;                                    // source samples per row (plane height H); 2 of them per 4 outputs
      const uint8_t* plane = aux + (size_t)pl * sw * a.H;
      for (int i = threadIdx.x; i < nwords; i += INGEST_THREADS) {
        const int y = i / q, x0 = (i - y * q) * 4, k0 = x0 >> 1;
        const uint8_t* row = plane + (size_t)y * sw;
        int p[5];                                                       // samples k0-1 .. k0+3, clamped
#pragma unroll
        for (int k = 0; k < 5; k++) p[k] = (int)__ldg(row + min(max(k0 - 1 + k, 0), sw - 1));
        int o[4];
#pragma unroll
        for (int j = 0; j < 2; j++) {                                   // [1 110 18 -1] and [-3 50 86 -5], first tap at k-1 (y4m_input.c:434-455)
          o[2 * j] = clamp255((p[j] + 110 * p[j + 1] + 18 * p[j + 2] - p[j + 3] + 64) >> 7);
          o[2 * j + 1] = clamp255((-3 * p[j] + 50 * p[j + 1] + 86 * p[j + 2] - 5 * p[j + 3] + 64) >> 7);
        }
        dst[i] = pack_u8x4(o);
      }
      break;
    }
    case P64B_CHROMA_444:
    case P64B_CHROMA_444ALPHA: {
      const uint32_t* plane = reinterpret_cast<const uint32_t*>(aux + (size_t)pl * a.W * a.H);   // the plane's first cw*ch bytes, linearly
      for (int i = threadIdx.x; i < nwords; i += INGEST_THREADS) dst[i] = __ldg(plane + i);
      break;
    }
    default:                                                            // mono
      for (int i = threadIdx.x; i < nwords; i += INGEST_THREADS) dst[i] = 0x80808080u;
  }
}

}  // namespace p64b
