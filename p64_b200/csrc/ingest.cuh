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

// horizontal [4 -17 114 35 -9 1]/128 at x of a row of n samples
__device__ __forceinline__ int hshift6(const uint8_t* __restrict__ r, int x, int n) {
  auto at = [&](int i) { return (int)__ldg(r + min(max(i, 0), n - 1)); };
  return clamp255((4 * at(x - 2) - 17 * at(x - 1) + 114 * at(x) + 35 * at(x + 1) - 9 * at(x + 2) + at(x + 3) + 64) >> 7);
}

__global__ void __launch_bounds__(256) ingest_chroma_kernel(const IngestArgs a) {
  const int cw = a.W >> 1, ch = a.H >> 1, q = cw >> 2;                  // 4 output samples per thread
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_stream = 2ll * ch * q;
  if (t >= per_stream * a.n_streams) return;
  const int s = (int)(t / per_stream);
  int r = (int)(t - (long long)s * per_stream);
  const int pl = r / (ch * q); r -= pl * ch * q;
  const int y = r / q, x0 = (r - y * q) * 4;
  const uint8_t* aux = a.aux + (size_t)s * a.aux_stride;
  int o[4];
  switch (a.chroma) {
    case P64B_CHROMA_420MPEG2:
    case P64B_CHROMA_422: {
      const int sh = a.chroma == P64B_CHROMA_422 ? a.H : ch;            // source plane height; only its first ch rows are read
      const uint8_t* row = aux + (size_t)pl * cw * sh + (size_t)y * cw;
#pragma unroll
      for (int j = 0; j < 4; j++) o[j] = hshift6(row, x0 + j, cw);
      break;
    }
    case P64B_CHROMA_420PALDV: {
      const uint8_t* plane = aux + (size_t)pl * cw * ch;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        auto tmp = [&](int yy) { return hshift6(plane + (size_t)min(max(yy, 0), ch - 1) * cw, x0 + j, cw); };
        const int x = x0 + j; (void)x;
        o[j] = pl == 0 ? clamp255((tmp(y - 3) - 9 * tmp(y - 2) + 35 * tmp(y - 1) + 114 * tmp(y) - 17 * tmp(y + 1) + 4 * tmp(y + 2) + 64) >> 7)
                       : clamp255((4 * tmp(y - 2) - 17 * tmp(y - 1) + 114 * tmp(y) + 35 * tmp(y + 1) - 9 * tmp(y + 2) + tmp(y + 3) + 64) >> 7);
      }
      break;
    }
    case P64B_CHROMA_411: {
      const int sw = (a.W + 3) >> 2;                                    // source samples per row; plane height H
      const uint8_t* row = aux + (size_t)pl * sw * a.H + (size_t)y * sw;
      auto at = [&](int i) { return (int)__ldg(row + min(max(i, 0), sw - 1)); };
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int k = (x0 + j) >> 1;
        o[j] = ((x0 + j) & 1) ? clamp255((-3 * at(k - 1) + 50 * at(k) + 86 * at(k + 1) - 5 * at(k + 2) + 64) >> 7)
                              : clamp255((at(k - 1) + 110 * at(k) + 18 * at(k + 1) - at(k + 2) + 64) >> 7);
      }
      break;
    }
    case P64B_CHROMA_444:
    case P64B_CHROMA_444ALPHA: {
      const uint8_t* p = aux + (size_t)pl * a.W * a.H + (size_t)y * cw + x0;   // the plane's first cw*ch bytes, linearly
#pragma unroll
      for (int j = 0; j < 4; j++) o[j] = __ldg(p + j);
      break;
    }
    default:                                                            // mono
      o[0] = o[1] = o[2] = o[3] = 128;
  }
  uint8_t* d = a.dst + (size_t)s * a.dst_stride + (size_t)a.W * a.H + (size_t)pl * cw * ch + (size_t)y * cw + x0;
  *reinterpret_cast<uint32_t*>(d) = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
}

}  // namespace p64b
