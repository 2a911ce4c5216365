// Device context and the p64b_ctx_* C ABI (include/p64_b200.h): frame stores resident in HBM for a batch
// of independent streams, kernel launches, host<->device staging.  No CPU fallback: every entry point
// fails with P64B_ECUDA when no CUDA device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "ingest.cuh"
#include "kernels.cuh"
#include "vlc_kernels.cuh"

namespace p64b {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                               \
      return P64B_ECUDA;                                                                           \
    }                                                                                              \
  } while (0)

static bool geom_for(int image_type, Geom* g) {
  switch (image_type) {                      // SetCCITT, p64.c:1476-1514
    case P64B_IT_NTSC: g->W = 352; g->H = 240; g->ngob = 10; g->qcif = 0; break;
    case P64B_IT_CIF:  g->W = 352; g->H = 288; g->ngob = 12; g->qcif = 0; break;
    case P64B_IT_QCIF: g->W = 176; g->H = 144; g->ngob = 3;  g->qcif = 1; break;
    default: return false;
  }
  g->mbw = g->W / 16; g->mbh = g->H / 16; g->nmb = g->ngob * 33; g->frame_bytes = g->W * g->H * 3 / 2;
  return true;
}

}  // namespace p64b

using namespace p64b;

struct p64b_ctx {
  int device = 0, image_type = 0, S = 0;
  Geom g{};
  cudaStream_t own = nullptr, stream = nullptr;
  uint8_t* d_src = nullptr;       // [S][frame_bytes]   (= slot 0 of the pipelined path)
  // pipelined host path (submit / wait): NSLOT sets of device staging buffers, separate copy streams
#ifndef P64B_NSLOT
#define P64B_NSLOT 3
#endif
  static constexpr int NSLOT = P64B_NSLOT;    // steps in flight (4 and 5 were measured: no gain)
  uint8_t* p_src[NSLOT] = {};
  p64b_mb* p_mbs[NSLOT] = {};
  int8_t* p_levels[NSLOT] = {};
  cudaEvent_t ev_h2d[NSLOT] = {}, ev_comp[NSLOT] = {}, ev_d2h[NSLOT] = {};
  bool slot_used[NSLOT] = {};
  bool slot_bits[NSLOT] = {};      // the slot's last step came from p64b_ctx_submit_bits
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_fix = nullptr;   // s_fix: the rare second download of a frame that outgrew its budget
  int64_t submitted = 0;
  uint8_t* d_fs[2] = {nullptr, nullptr};   // frame stores; d_fs[cur] = CFS (reference), d_fs[cur^1] = OFS
  uint8_t* d_li[2] = {nullptr, nullptr};   // LastIntra, same double buffering
  int cur = 0;
  p64b_me* d_me = nullptr;        // [S][nmb]
  p64b_mb* d_mbs = nullptr;       // [S][nmb]
  int8_t* d_levels = nullptr;     // [S][nmb][384]
  uint8_t* d_quant = nullptr;     // [S]
  uint8_t* d_ovf = nullptr;       // [S][nmb]
  const uint8_t* frame_src = nullptr;   // source of the frame in flight (frame_begin .. frame_end)
  const uint8_t* last_src = nullptr;    // device source frames of the last coded frame (for p64b_ctx_statistics)
  p64b_plane_stats* d_stats = nullptr;  // [S][3], allocated on first use
  int64_t launches = 0;
  bool me_attr_done = false, mb_attr_done = false, dec_attr_done = false;   // per context: function attributes are per device
  int n_sm = 0;
  uint32_t* d_me_queue = nullptr;   // [2] work counters of the persistent ME kernel (alternating per launch)
  int64_t me_launches = 0, vlc_launches = 0;
  int vlc_ctas_per_sm = 0;
  // device-side entropy coding (p64b_ctx_submit_bits / p64b_ctx_wait_bits), allocated on first use
  DevVlcTables* d_vlc_tables = nullptr;
  uint32_t* d_gob_words = nullptr;          // [S][ngob][VLC_GOB_WORDS] GOB bit strings
  uint32_t* d_gob_bits = nullptr;           // [S][ngob]
  uint32_t *d_carry = nullptr, *d_carry_len = nullptr;     // [S] pending bits (< 8) of every stream
  unsigned long long* d_bitpos = nullptr;   // [S] bits written so far
  size_t bits_out_cap = 0;                  // bytes of one output buffer (worst case)
  uint8_t* d_bits_out[NSLOT] = {};
  uint8_t* h_bits_out[NSLOT] = {};          // pinned; grown on demand
  size_t h_bits_cap[NSLOT] = {}, slot_copied[NSLOT] = {};
  size_t bits_budget = 0;                   // data bytes downloaded with the first copy (adapts to the last frames' sizes)
  size_t bits_recent[4] = {};               // totals of the last four steps collected
  int64_t second_copies = 0;                // steps whose frame outgrew the budget (completed by a second, synchronous copy)
  int pending_d2h = -1;                     // slot whose download waits for the next step's upload to finish (see p64b_ctx_submit_bits)
  // ingest (p64b_ctx_set_input_chroma): host sources are unconverted Y4M payloads; chroma converted on the device
  int chroma = P64B_CHROMA_420JPEG;
  size_t raw_bytes = 0, aux_bytes = 0;      // per frame: whole payload; the part of its chroma the conversion reads
  uint8_t* d_aux[NSLOT] = {};               // [S][aux_bytes] uploaded chroma payload per pipeline slot
  // rate control on the device (p64b_ctx_set_rate_control)
  p64b_rate_control rate{};                 // rate.rate == 0: off
  uint32_t* d_frame_bits = nullptr;         // [S] bits of the frame in flight so far
  long long* d_buffer_offset = nullptr;     // [S] BufferOffset
  uint32_t* d_overflows = nullptr;          // [S] NumberOvfl
  bool rc_started = false;                  // the initial quantiser has been loaded
  uint32_t* d_pic_hdr = nullptr;            // [NSLOT][2] picture header of the step in each pipeline slot
  // The rate-control frame step is a launch-bound chain of 1 + 2 x NumberGOB + 5 small kernels: it is captured once per
  // (pipeline slot, frame-store parity, ME queue parity, search) into a CUDA graph and replayed (one launch per frame).
  struct RcGraph { int slot, cur, parity, me_mode, search_limit, force_intra; cudaGraphExec_t exec; int kernels; };
  std::vector<RcGraph> rc_graphs;
  bool use_graphs = true;
  // optional per-kernel timing with CUDA events on the launching stream (bench.py roofline)
  bool prof = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[3];   // 0 = ME kernel, 1 = MB kernel, 2 = entropy-coding kernels
};

struct ProfScope {   // records an event pair around one launch when profiling is on
  p64b_ctx* c; int k; cudaEvent_t e1 = nullptr;
  ProfScope(p64b_ctx* c_, int k_) : c(c_), k(k_) {
    if (!c->prof) return;
    cudaEvent_t e0;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e1 = nullptr; return; }
    cudaEventRecord(e0, c->stream);
    c->prof_ev[k].emplace_back(e0, e1);
  }
  ~ProfScope() { if (e1) cudaEventRecord(e1, c->stream); }
};

static int use_device(const p64b_ctx* c) {
  CU(cudaSetDevice(c->device));
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// u8 tensor {W, H, n_pairs} over luma planes `stride` bytes apart; box {bw, bh, 1}; out-of-bounds -> 0
static int make_plane_map(CUtensorMap* tm, const uint8_t* base, int W, int H, size_t stride, int n_pairs, int bw, int bh) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return P64B_ECUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride & 15)) { set_error("luma planes must be 16-byte aligned"); return P64B_EINVAL; }
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_pairs};
  cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)stride};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return P64B_ECUDA; }
  return 0;
}

// Legal surface positions (d+15) per macroblock column / row class, first | interior | last: inside the frame with the
// reference's strict < on the far edge (me.c:212-213, 292-293) and inside the search range -- FastBME scans
// [-S/2, S/2) (me.c:206-208), StepBME reaches [-15,15] (me.c:294).  Interior macroblocks see the whole range
// (16 > 15).  One byte per class; hi is stored +1 so that 0 means "empty".
static MeRanges me_ranges(const Geom& g, int me_mode, int search_limit) {
  const bool full = me_mode == P64B_ME_FULL;
  const int rlo = full ? 15 - search_limit / 2 : 0, rhi = full ? 14 + search_limit / 2 : 30;
  MeRanges r{0, 0, 0, 0};
  for (int cls = 0; cls < 3; cls++) {
    const int x0 = cls == 0 ? 0 : (cls == 1 ? 16 : g.W - 16), y0 = cls == 0 ? 0 : (cls == 1 ? 16 : g.H - 16);
    int xlo = std::max(rlo, 15 - x0), xhi = std::min(rhi, g.W - 2 - x0);
    int ylo = std::max(rlo, 15 - y0), yhi = std::min(rhi, g.H - 2 - y0);
    if (xhi < xlo) { xlo = 15; xhi = -1; }
    if (yhi < ylo) { ylo = 15; yhi = -1; }
    r.xlo |= (uint32_t)xlo << (8 * cls); r.xhi |= (uint32_t)(xhi + 1) << (8 * cls);
    r.ylo |= (uint32_t)ylo << (8 * cls); r.yhi |= (uint32_t)(yhi + 1) << (8 * cls);
  }
  return r;
}

static int launch_me(p64b_ctx* c, const uint8_t* ref, const uint8_t* cur, size_t stride, int n_pairs, int me_mode,
                     int search_limit, p64b_me* out, uint32_t* surface = nullptr) {
  int rc;
  if (!c->me_attr_done) {
    CU(cudaFuncSetAttribute(me_search_kernel<ME_V_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_FULL));
    CU(cudaFuncSetAttribute(me_search_kernel<ME_V_SURF>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_SURF));
    CU(cudaFuncSetAttribute(me_search_kernel<ME_V_TSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_FULL));
    CU(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, c->device));
    c->me_attr_done = true;
  }
  CUtensorMap tm_ref, tm_cur;
  if ((rc = make_plane_map(&tm_ref, ref, c->g.W, c->g.H, stride, n_pairs, 48, ME_WIN_ROWS))) return rc;
  if ((rc = make_plane_map(&tm_cur, cur, c->g.W, c->g.H, stride, n_pairs, 16, 16))) return rc;
  const bool surf = surface != nullptr;
  MeArgs a;
  a.rg = me_ranges(c->g, me_mode, search_limit);
  a.me_mode = me_mode; a.mbw = c->g.mbw; a.mbh = c->g.mbh; a.n_pairs = n_pairs; a.out = out; a.surface = surface;
  // persistent worker warps: as many CTAs as stay resident (4 per SM; 3 with the surface in shared memory)
  const long long total = (long long)n_pairs * a.mbw * a.mbh;
  if (total > 0x7fffffffLL) { set_error("too many macroblocks in one motion-estimation call"); return P64B_EINVAL; }
  const int grid = (int)std::min<long long>((long long)c->n_sm * (surf ? 3 : 4), (total + ME_WARPS - 1) / ME_WARPS);
  // n / per_pair by multiply-high (exact for n < 2^31: the magic's excess e < per_pair and n * e < 2^(32+shift))
  const uint32_t per_pair = (uint32_t)(a.mbw * a.mbh);
  uint32_t sh = 0;
  while ((2u << sh) < per_pair) sh++;                       // 2^sh < per_pair <= 2^(sh+1): the magic fits 32 bits
  a.shift_pp = sh;
  a.magic_pp = (uint32_t)(((1ull << (32 + sh)) + per_pair - 1) / per_pair);
  a.magic_w = 65536u / (uint32_t)a.mbw + 1u;               // r / mbw for r < per_pair <= 1024
  a.queue = c->d_me_queue; a.parity = (int)(c->me_launches++ & 1);
  a.executed = reinterpret_cast<unsigned long long*>(c->d_me_queue + 2);
  a.m8 = 1u << 8; a.m16 = 1u << 16; a.m24 = 1u << 24; a.m2048 = 2048u;
  ProfScope ps(c, 0);
  if (surf)                          me_search_kernel<ME_V_SURF><<<grid, ME_THREADS, ME_SMEM_SURF, c->stream>>>(tm_ref, tm_cur, a);
  else if (me_mode == P64B_ME_FULL)  me_search_kernel<ME_V_FULL><<<grid, ME_THREADS, ME_SMEM_FULL, c->stream>>>(tm_ref, tm_cur, a);
  else                               me_search_kernel<ME_V_TSS><<<grid, ME_THREADS, ME_SMEM_FULL, c->stream>>>(tm_ref, tm_cur, a);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

// frame_layout: the outputs of a partial launch land at their position in the whole-frame arrays [S][nmb] instead of
// being packed [S][gob_count*33]
static int launch_mb(p64b_ctx* c, const p64b_step* st, const uint8_t* src, int gob_first, int gob_count,
                     const uint8_t* d_quant, p64b_mb* mbs, int8_t* levels, bool frame_layout = false) {
  MbArgs a;
  a.g = c->g; a.src = src; a.ref = c->d_fs[c->cur]; a.out = c->d_fs[c->cur ^ 1];
  a.me = c->d_me; a.li_prev = c->d_li[c->cur]; a.li_new = c->d_li[c->cur ^ 1];
  a.quant = d_quant; a.mbs = mbs; a.levels = levels; a.n_streams = c->S;
  a.gob_first = gob_first; a.gob_count = gob_count;
  a.out_mb_per_stream = frame_layout ? c->g.nmb : gob_count * 33;
  if (frame_layout) { a.mbs += gob_first * 33; a.levels += (size_t)gob_first * 33 * P64B_LEVELS_PER_MB; }
  a.first_frame = st->first_frame; a.force_intra = st->force_intra; a.gquant = st->gquant;
  const int n = c->S * gob_count * 33;
  {   // n / mps by multiply-high, exact for n < 2^31 (same construction as in launch_me)
    const uint32_t mps = (uint32_t)gob_count * 33u;
    uint32_t sh = 0;
    while ((2u << sh) < mps) sh++;
    a.mps_shift = sh; a.mps_magic = (uint32_t)(((1ull << (32 + sh)) + mps - 1) / mps);
  }
  ProfScope ps(c, 1);
  if (!c->mb_attr_done) {
    CU(cudaFuncSetAttribute(mb_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB4_SMEM));
    c->mb_attr_done = true;
  }
  mb_encode_kernel<<<(n + MB4_PER_CTA - 1) / MB4_PER_CTA, MB4_THREADS, MB4_SMEM, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

static int check_step(const p64b_step* st) {
  if (!st) { set_error("step is NULL"); return P64B_EINVAL; }
  if (st->me_mode != P64B_ME_TSS && st->me_mode != P64B_ME_FULL) { set_error("bad me_mode"); return P64B_EINVAL; }
  if (st->search_limit < 1 || st->search_limit > 31) { set_error("search_limit out of [1,31]"); return P64B_EINVAL; }
  return 0;
}
static int check_quant(int q) {
  if (q < 1 || q > 31) { set_error("quantiser out of [1,31]"); return P64B_EINVAL; }
  return 0;
}

extern "C" {

const char* p64b_last_error(void) { return g_err.c_str(); }
int p64b_version(void) { return P64B_VERSION; }

int p64b_width(int t) { Geom g; return geom_for(t, &g) ? g.W : P64B_EINVAL; }
int p64b_height(int t) { Geom g; return geom_for(t, &g) ? g.H : P64B_EINVAL; }
int p64b_frame_bytes(int t) { Geom g; return geom_for(t, &g) ? g.frame_bytes : P64B_EINVAL; }
int p64b_num_gob(int t) { Geom g; return geom_for(t, &g) ? g.ngob : P64B_EINVAL; }
int p64b_num_mb(int t) { Geom g; return geom_for(t, &g) ? g.nmb : P64B_EINVAL; }

int p64b_ctx_create(p64b_ctx** out, int device, int image_type, int n_streams) {
  if (!out || n_streams < 1) { set_error("bad arguments"); return P64B_EINVAL; }
  Geom g;
  if (!geom_for(image_type, &g)) { set_error("unknown image type"); return P64B_EINVAL; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (p64_b200 has no CPU fallback)");
    return P64B_ECUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return P64B_EINVAL; }
  p64b_ctx* c = new p64b_ctx();
  c->device = device; c->image_type = image_type; c->S = n_streams; c->g = g;
  auto fail = [&](int rc) { p64b_ctx_destroy(c); return rc; };
  if (use_device(c)) return fail(P64B_ECUDA);
  const size_t fb = (size_t)n_streams * g.frame_bytes, slack = 64, nm = (size_t)n_streams * g.nmb;
#define ALLOC(p, bytes)                                                                                  \
  if (cudaMalloc((void**)&(p), (bytes)) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(P64B_ENOMEM); } \
  if (cudaMemset((p), 0, (bytes)) != cudaSuccess) { set_error("cudaMemset failed"); return fail(P64B_ECUDA); }
  ALLOC(c->d_src, fb + slack);
  ALLOC(c->d_fs[0], fb + slack);           // ClearFS: both stores start zero-filled (p64.c:532-535)
  ALLOC(c->d_fs[1], fb + slack);
  ALLOC(c->d_li[0], nm);                   // LastIntra zeroed (p64.c:1522-1528)
  ALLOC(c->d_li[1], nm);
  ALLOC(c->d_me, nm * sizeof(p64b_me));    // Me* arrays are zero on the first frame (me.c:49-59)
  ALLOC(c->d_mbs, nm * sizeof(p64b_mb));
  ALLOC(c->d_levels, nm * P64B_LEVELS_PER_MB);
  ALLOC(c->d_quant, (size_t)n_streams);
  ALLOC(c->d_ovf, nm);
  ALLOC(c->d_me_queue, 64 * sizeof(uint32_t));  // two work counters + the 64-bit executed-work counter (+ room: the kernel forms one
                                                // address per lane for its predicated queue atomic, only lane 0's is ever used)
  c->p_src[0] = c->d_src; c->p_mbs[0] = c->d_mbs; c->p_levels[0] = c->d_levels;
  for (int i = 1; i < p64b_ctx::NSLOT; i++) {
    ALLOC(c->p_src[i], fb + slack);
    ALLOC(c->p_mbs[i], nm * sizeof(p64b_mb));
    ALLOC(c->p_levels[i], nm * P64B_LEVELS_PER_MB);
  }
  for (int i = 0; i < p64b_ctx::NSLOT; i++)
    if (cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming) != cudaSuccess) { set_error("event create failed"); return fail(P64B_ECUDA); }
  if (cudaStreamCreateWithFlags(&c->own, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->s_fix, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); return fail(P64B_ECUDA); }
  c->stream = c->own;
#undef ALLOC
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("sync failed"); return fail(P64B_ECUDA); }
  *out = c;
  return 0;
}

void p64b_ctx_destroy(p64b_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_src); cudaFree(c->d_fs[0]); cudaFree(c->d_fs[1]); cudaFree(c->d_li[0]); cudaFree(c->d_li[1]);
  cudaFree(c->d_me); cudaFree(c->d_mbs); cudaFree(c->d_levels); cudaFree(c->d_quant); cudaFree(c->d_ovf); cudaFree(c->d_me_queue);
  if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
  if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
  for (int i = 0; i < p64b_ctx::NSLOT; i++) {
    if (i) { cudaFree(c->p_src[i]); cudaFree(c->p_mbs[i]); cudaFree(c->p_levels[i]); }
    if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
    if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
    if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
  }
  cudaFree(c->d_stats);
  cudaFree(c->d_vlc_tables); cudaFree(c->d_gob_words); cudaFree(c->d_gob_bits); cudaFree(c->d_carry); cudaFree(c->d_carry_len);
  for (int i = 0; i < p64b_ctx::NSLOT; i++) cudaFree(c->d_aux[i]);
  for (auto& g : c->rc_graphs) cudaGraphExecDestroy(g.exec);
  cudaFree(c->d_pic_hdr);
  cudaFree(c->d_bitpos); cudaFree(c->d_frame_bits); cudaFree(c->d_buffer_offset); cudaFree(c->d_overflows);
  for (int i = 0; i < p64b_ctx::NSLOT; i++) { cudaFree(c->d_bits_out[i]); if (c->h_bits_out[i]) cudaFreeHost(c->h_bits_out[i]); }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->s_fix) cudaStreamDestroy(c->s_fix);
  if (c->own) cudaStreamDestroy(c->own);
  delete c;
}

int p64b_ctx_streams(const p64b_ctx* c) { return c ? c->S : P64B_EINVAL; }

int p64b_ctx_set_cuda_stream(p64b_ctx* c, void* s) {
  if (!c) return P64B_EINVAL;
  c->stream = s ? (cudaStream_t)s : c->own;
  return 0;
}

void* p64b_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc failed"); return nullptr; }   // (write-combined: measured, no gain: 55.2 GB/s either way)
  return p;
}
void* p64b_host_alloc_flags(size_t bytes, int flags) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, (flags & 1) ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc failed"); return nullptr; }
  return p;
}
void p64b_host_free(void* p) { if (p) cudaFreeHost(p); }

static void swap_stores(p64b_ctx* c) { c->cur ^= 1; }    // SwapFS(CFS,OFS), p64.c:661

// H2D of one step's host source frames into `dst` (a staging slot) on `copy_stream`.  420jpeg: one contiguous copy.
// Otherwise the luma planes go straight into the frames and the chroma payload into the slot's aux buffer (two strided
// copies); ingest_source() then converts it on the compute stream.
static int upload_source(p64b_ctx* c, int slot, uint8_t* dst, const uint8_t* src, cudaStream_t copy_stream) {
  const size_t fb = (size_t)c->g.frame_bytes, wh = (size_t)c->g.W * c->g.H;
  if (c->chroma == P64B_CHROMA_420JPEG) {
    CU(cudaMemcpyAsync(dst, src, (size_t)c->S * fb, cudaMemcpyHostToDevice, copy_stream));
    return 0;
  }
  CU(cudaMemcpy2DAsync(dst, fb, src, c->raw_bytes, wh, (size_t)c->S, cudaMemcpyHostToDevice, copy_stream));
  if (c->aux_bytes)
    CU(cudaMemcpy2DAsync(c->d_aux[slot], c->aux_bytes, src + wh, c->raw_bytes, c->aux_bytes, (size_t)c->S, cudaMemcpyHostToDevice, copy_stream));
  return 0;
}
static int ingest_source(p64b_ctx* c, int slot, uint8_t* dst) {
  if (c->chroma == P64B_CHROMA_420JPEG) return 0;
  IngestArgs a;
  a.aux = c->d_aux[slot]; a.aux_stride = c->aux_bytes; a.dst = dst; a.dst_stride = (size_t)c->g.frame_bytes;
  a.W = c->g.W; a.H = c->g.H; a.n_streams = c->S; a.chroma = c->chroma;
  const size_t smem = c->chroma == P64B_CHROMA_420PALDV ? (size_t)(c->g.W / 2) * (c->g.H / 2) : 0;   // the intermediate plane
  ingest_chroma_kernel<<<c->S * 2, INGEST_THREADS, smem, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

int p64b_ctx_encode_frames_dev(p64b_ctx* c, const p64b_step* st, const uint8_t* src_dev, p64b_mb* mbs_dev,
                               int8_t* levels_dev) {
  if (!c || !src_dev || !mbs_dev || !levels_dev) { set_error("NULL argument"); return P64B_EINVAL; }
  int rc;
  if ((rc = check_step(st)) || (rc = check_quant(st->gquant)) || (rc = use_device(c))) return rc;
  if (!st->first_frame) {
    if ((rc = launch_me(c, c->d_fs[c->cur], src_dev, (size_t)c->g.frame_bytes, c->S, st->me_mode, st->search_limit, c->d_me))) return rc;
  } else {
    CU(cudaMemsetAsync(c->d_me, 0, (size_t)c->S * c->g.nmb * sizeof(p64b_me), c->stream));
  }
  if ((rc = launch_mb(c, st, src_dev, 0, c->g.ngob, nullptr, mbs_dev, levels_dev))) return rc;
  swap_stores(c);
  c->last_src = src_dev;
  return 0;
}

// Pipelined host path.  submit(n) enqueues, without blocking the host:
//   copy stream  : H2D of the step's source frames into staging slot n % NSLOT
//   compute stream (the context's stream): ME + MB kernels (ordered after the H2D; frame stores chain the steps)
//   copy-back stream : D2H of records + levels into the caller's buffers
// so that step n+1's upload and step n-1's download overlap step n's kernels.  wait(ticket) blocks until the
// caller's output buffers of that step are complete.  Host buffers should be pinned (p64b_host_alloc).
int p64b_ctx_submit(p64b_ctx* c, const p64b_step* st, const uint8_t* src, p64b_mb* mbs, int8_t* levels, int64_t* ticket) {
  if (!c || !src || !mbs || !levels || !ticket) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->frame_src) { set_error("p64b_ctx_submit inside frame_begin/frame_end"); return P64B_EINVAL; }
  int rc;
  if ((rc = check_step(st)) || (rc = check_quant(st->gquant)) || (rc = use_device(c))) return rc;
  const int slot = (int)(c->submitted % p64b_ctx::NSLOT);
  const size_t nm = (size_t)c->S * c->g.nmb;
  if (c->slot_used[slot]) {
    CU(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[slot], 0));    // the slot's source was consumed by its last kernels
    CU(cudaStreamWaitEvent(c->stream, c->ev_d2h[slot], 0));    // its outputs were downloaded
  }
  if ((rc = upload_source(c, slot, c->p_src[slot], src, c->s_h2d))) return rc;
  CU(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
  CU(cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0));
  if ((rc = ingest_source(c, slot, c->p_src[slot]))) return rc;
  if ((rc = p64b_ctx_encode_frames_dev(c, st, c->p_src[slot], c->p_mbs[slot], c->p_levels[slot]))) return rc;
  CU(cudaEventRecord(c->ev_comp[slot], c->stream));
  CU(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[slot], 0));
  CU(cudaMemcpyAsync(mbs, c->p_mbs[slot], nm * sizeof(p64b_mb), cudaMemcpyDeviceToHost, c->s_d2h));
  CU(cudaMemcpyAsync(levels, c->p_levels[slot], nm * P64B_LEVELS_PER_MB, cudaMemcpyDeviceToHost, c->s_d2h));
  CU(cudaEventRecord(c->ev_d2h[slot], c->s_d2h));
  c->slot_used[slot] = true;
  c->slot_bits[slot] = false;
  *ticket = c->submitted++;
  return 0;
}

int p64b_ctx_wait(p64b_ctx* c, int64_t ticket) {
  if (!c || ticket < 0 || ticket >= c->submitted) { set_error("bad ticket"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaEventSynchronize(c->ev_d2h[ticket % p64b_ctx::NSLOT]));
  return 0;
}

int p64b_ctx_encode_frames(p64b_ctx* c, const p64b_step* st, const uint8_t* src, p64b_mb* mbs, int8_t* levels) {
  int64_t t;
  int rc = p64b_ctx_submit(c, st, src, mbs, levels, &t);
  return rc ? rc : p64b_ctx_wait(c, t);
}

int p64b_ctx_frame_begin(p64b_ctx* c, const p64b_step* st, const uint8_t* src) {
  if (!c || !src) { set_error("NULL argument"); return P64B_EINVAL; }
  int rc;
  if ((rc = check_step(st)) || (rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->s_h2d));      // drain any pipelined steps still in flight (they share staging slot 0)
  CU(cudaStreamSynchronize(c->s_d2h));
  if ((rc = upload_source(c, 0, c->d_src, src, c->stream)) || (rc = ingest_source(c, 0, c->d_src))) return rc;
  if (!st->first_frame) {
    if ((rc = launch_me(c, c->d_fs[c->cur], c->d_src, (size_t)c->g.frame_bytes, c->S, st->me_mode, st->search_limit, c->d_me))) return rc;
  } else {
    CU(cudaMemsetAsync(c->d_me, 0, (size_t)c->S * c->g.nmb * sizeof(p64b_me), c->stream));
  }
  c->frame_src = c->d_src;
  return 0;
}

int p64b_ctx_encode_gob(p64b_ctx* c, const p64b_step* st, int gob, const uint8_t* quant, p64b_mb* mbs, int8_t* levels) {
  if (!c || !quant || !mbs || !levels) { set_error("NULL argument"); return P64B_EINVAL; }
  if (!c->frame_src) { set_error("p64b_ctx_encode_gob outside frame_begin/frame_end"); return P64B_EINVAL; }
  if (gob < 0 || gob >= c->g.ngob) { set_error("gob out of range"); return P64B_EINVAL; }
  int rc;
  if ((rc = check_step(st)) || (rc = use_device(c))) return rc;
  for (int s = 0; s < c->S; s++) if ((rc = check_quant(quant[s]))) return rc;
  CU(cudaMemcpyAsync(c->d_quant, quant, (size_t)c->S, cudaMemcpyHostToDevice, c->stream));
  // outputs of one GOB are packed [S][33] at the head of the context's output buffers
  if ((rc = launch_mb(c, st, c->frame_src, gob, 1, c->d_quant, c->d_mbs, c->d_levels))) return rc;
  const size_t nm = (size_t)c->S * 33;
  CU(cudaMemcpyAsync(mbs, c->d_mbs, nm * sizeof(p64b_mb), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(levels, c->d_levels, nm * P64B_LEVELS_PER_MB, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int p64b_ctx_frame_end(p64b_ctx* c, const uint8_t* overflow) {
  if (!c) return P64B_EINVAL;
  if (!c->frame_src) { set_error("p64b_ctx_frame_end without frame_begin"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  if (overflow) {
    const size_t nm = (size_t)c->S * c->g.nmb;
    bool any = false;
    for (size_t i = 0; i < nm && !any; i++) any = overflow[i] != 0;
    if (any) {
      CU(cudaMemcpyAsync(c->d_ovf, overflow, nm, cudaMemcpyHostToDevice, c->stream));
      const int warps = (int)nm, threads = 256;
      overflow_patch_kernel<<<(warps * 32 + threads - 1) / threads, threads, 0, c->stream>>>(
          c->g, c->d_ovf, c->d_fs[c->cur], c->d_fs[c->cur ^ 1], c->d_li[c->cur], c->d_li[c->cur ^ 1], c->S);
      c->launches++;
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(c->stream));   // `overflow` is caller memory
    }
  }
  swap_stores(c);
  c->last_src = c->frame_src;
  c->frame_src = nullptr;
  return 0;
}

int p64b_ctx_motion_estimation_dev(p64b_ctx* c, const uint8_t* ref_dev, const uint8_t* cur_dev, int n_pairs,
                                   int me_mode, int search_limit, p64b_me* out_dev) {
  if (!c || !ref_dev || !cur_dev || !out_dev || n_pairs < 1) { set_error("bad arguments"); return P64B_EINVAL; }
  p64b_step st{}; st.me_mode = me_mode; st.search_limit = search_limit;
  int rc;
  if ((rc = check_step(&st)) || (rc = use_device(c))) return rc;
  return launch_me(c, ref_dev, cur_dev, (size_t)c->g.W * c->g.H, n_pairs, me_mode, search_limit, out_dev);
}

int p64b_ctx_sad_surface_dev(p64b_ctx* c, const uint8_t* ref_dev, const uint8_t* cur_dev, int n_pairs, p64b_me* out_dev,
                             uint32_t* surface_dev) {
  if (!c || !ref_dev || !cur_dev || !out_dev || !surface_dev || n_pairs < 1) { set_error("bad arguments"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  return launch_me(c, ref_dev, cur_dev, (size_t)c->g.W * c->g.H, n_pairs, P64B_ME_TSS, 31, out_dev, surface_dev);
}

int p64b_ctx_me_records(p64b_ctx* c, int stream, p64b_me* out) {
  if (!c || !out || stream < 0 || stream >= c->S) { set_error("bad arguments"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(out, c->d_me + (size_t)stream * c->g.nmb, (size_t)c->g.mbw * c->g.mbh * sizeof(p64b_me), cudaMemcpyDeviceToHost));
  return 0;
}

int p64b_ctx_download_recon(p64b_ctx* c, int stream, uint8_t* yuv) {
  if (!c || !yuv || stream < 0 || stream >= c->S) { set_error("bad arguments"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(yuv, c->d_fs[c->cur] + (size_t)stream * c->g.frame_bytes, c->g.frame_bytes, cudaMemcpyDeviceToHost));
  return 0;
}

// Decoder's inverse half for one picture of every stream: records + levels (from the bit-stream parser) up, mb_decode_kernel,
// SwapFS.  The decoded picture is then the context's reference store (p64b_ctx_download_recon).
int p64b_ctx_decode_frames(p64b_ctx* c, const p64b_mb* mbs, const int8_t* levels) {
  if (!c || !mbs || !levels) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->frame_src) { set_error("p64b_ctx_decode_frames inside frame_begin/frame_end"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->s_h2d));      // slot 0 buffers may still be in use by pipelined encode steps
  CU(cudaStreamSynchronize(c->s_d2h));
  const size_t nm = (size_t)c->S * c->g.nmb;
  CU(cudaMemcpyAsync(c->d_mbs, mbs, nm * sizeof(p64b_mb), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_levels, levels, nm * P64B_LEVELS_PER_MB, cudaMemcpyHostToDevice, c->stream));
  MbDecArgs a;
  a.g = c->g; a.ref = c->d_fs[c->cur]; a.out = c->d_fs[c->cur ^ 1]; a.mbs = c->d_mbs; a.levels = c->d_levels; a.n_streams = c->S;
  if (!c->dec_attr_done) { CU(cudaFuncSetAttribute(mb_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB4_SMEM)); c->dec_attr_done = true; }
  mb_decode_kernel<<<((int)nm + MB4_PER_CTA - 1) / MB4_PER_CTA, MB4_THREADS, MB4_SMEM, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));     // mbs / levels are caller memory
  swap_stores(c);
  c->last_src = nullptr;
  return 0;
}

int p64b_ctx_statistics(p64b_ctx* c, p64b_plane_stats* out) {
  if (!c || !out) { set_error("NULL argument"); return P64B_EINVAL; }
  if (!c->last_src || c->frame_src) { set_error("p64b_ctx_statistics needs a completely coded frame"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  const size_t bytes = (size_t)c->S * 3 * sizeof(p64b_plane_stats);
  if (!c->d_stats && cudaMalloc((void**)&c->d_stats, bytes) != cudaSuccess) { set_error("cudaMalloc failed (statistics)"); return P64B_ENOMEM; }
  CU(cudaMemsetAsync(c->d_stats, 0, bytes, c->stream));
  const int wh = c->g.W * c->g.H;
  const int per_stream = (wh + STATS_CHUNK - 1) / STATS_CHUNK + 2 * ((wh / 4 + STATS_CHUNK - 1) / STATS_CHUNK);
  plane_stats_kernel<<<c->S * per_stream, 256, 0, c->stream>>>(c->g, c->last_src, c->d_fs[c->cur], c->d_stats, c->S);
  c->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->d_stats, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// StatisticsMem's derived quantities (stat.c:104-129) from the integer sums; the reference accumulates the same
// integers in double, which is exact below 2^53.
void p64b_stat_from_sums(const p64b_plane_stats* m, p64b_stat* s) {
  if (!m || !s) return;
  const double top = (double)m->n, value = (double)m->sum_rec, squared = (double)m->sum_sq_err;
  const double rsquared = (double)m->sum_sq_src, rvalue = (double)m->sum_src;
  s->mean = value / top;
  s->mse = squared / top;
  if (squared) {
    s->snr = rsquared ? 10 * log10(rsquared / squared) : -99.99;
    const double mr = rsquared - (rvalue * rvalue / top);
    s->mrsnr = mr ? 10 * log10(mr / squared) : -99.99;
    s->psnr = m->n ? 10 * log10((65025.0 * top) / squared) : -99.99;
  } else {
    s->snr = s->mrsnr = s->psnr = 99.99;
  }
  double e = 0;
  for (int i = 0; i < 256; i++)
    if (m->hist[i]) { const double p = (double)m->hist[i] / top; e += p * log(p); }
  s->entropy = -e / log(2.0);
}

int p64b_ctx_last_intra(p64b_ctx* c, int stream, uint8_t* out) {
  if (!c || !out || stream < 0 || stream >= c->S) { set_error("bad arguments"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(out, c->d_li[c->cur] + (size_t)stream * c->g.nmb, c->g.nmb, cudaMemcpyDeviceToHost));
  return 0;
}

int64_t p64b_ctx_launches(const p64b_ctx* c) { return c ? c->launches : 0; }
int64_t p64b_ctx_second_copies(const p64b_ctx* c) { return c ? c->second_copies : 0; }

int p64b_ctx_me_executed(p64b_ctx* c, uint64_t* packed_sad_ops, int reset) {
  if (!c || !packed_sad_ops) { set_error("NULL argument"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  unsigned long long rows = 0;
  CU(cudaMemcpy(&rows, c->d_me_queue + 2, sizeof rows, cudaMemcpyDeviceToHost));
  if (reset) CU(cudaMemset(c->d_me_queue + 2, 0, sizeof rows));
  *packed_sad_ops = (uint64_t)rows * 32u * 4u;      // a candidate row = 4 packed SADs in each of the warp's 32 lanes
  return 0;
}

int p64b_ctx_profile(p64b_ctx* c, int enable) {
  if (!c) return P64B_EINVAL;
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  for (auto& v : c->prof_ev) { for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } v.clear(); }
  c->prof = enable != 0;
  return 0;
}

int p64b_ctx_profile_read(p64b_ctx* c, double* ms_total, int32_t* count) {
  if (!c || !ms_total || !count) return P64B_EINVAL;
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; k++) {
    double t = 0;
    for (auto& e : c->prof_ev[k]) { float ms = 0; CU(cudaEventElapsedTime(&ms, e.first, e.second)); t += ms; }
    ms_total[k] = t; count[k] = (int32_t)c->prof_ev[k].size();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Device-side entropy coding: the frame step that returns bytes
// ---------------------------------------------------------------------------------------------------
static void free_bits_buffers(p64b_ctx* c) {
  void** ptrs[] = {(void**)&c->d_gob_words, (void**)&c->d_gob_bits, (void**)&c->d_carry, (void**)&c->d_carry_len, (void**)&c->d_bitpos,
                   (void**)&c->d_frame_bits, (void**)&c->d_buffer_offset, (void**)&c->d_overflows, (void**)&c->d_pic_hdr, (void**)&c->d_vlc_tables};
  for (void** p : ptrs) { cudaFree(*p); *p = nullptr; }
  for (int i = 0; i < p64b_ctx::NSLOT; i++) { cudaFree(c->d_bits_out[i]); c->d_bits_out[i] = nullptr; }
}

// Launch shape of the fixed-quantiser entropy kernel: as many CTAs as are resident at once (the warps take pieces from a queue),
// and this launch's queue counter (the kernel zeroes the other one for the next launch, like the motion search's).
static int vlc_seq_grid(p64b_ctx* c, VlcArgs* a, int* grid) {
  if (!c->vlc_ctas_per_sm) {
    if (!c->n_sm) CU(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, c->device));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->vlc_ctas_per_sm, vlc_gob_seq_kernel, VLC_SEQ_THREADS, 0));
    if (c->vlc_ctas_per_sm < 1) c->vlc_ctas_per_sm = 1;
  }
  const int tasks = c->S * c->g.ngob * VLC_PPG;
  *grid = std::min((tasks + VLC_SEQ_WARPS - 1) / VLC_SEQ_WARPS, c->n_sm * c->vlc_ctas_per_sm);
  a->queue = c->d_me_queue + 8; a->parity = (int)(c->vlc_launches++ & 1);
  return 0;
}

static int ensure_bits_buffers(p64b_ctx* c) {
  if (c->d_vlc_tables) return 0;          // allocated last: set only when every buffer exists (a failure below frees them all)
  const size_t S = (size_t)c->S, ng = (size_t)c->g.ngob;
  DevVlcTables t;
  fill_dev_vlc_tables(&t);
  // worst case of one frame: every block escapes every coefficient (GOB slot size) + picture header + carry
  c->bits_out_cap = vlc_data_offset(c->S) + S * (ng * VLC_GOB_WORDS * 4 + 16);
#define ALLOCZ(p, bytes)                                                                                  \
  if (cudaMalloc((void**)&(p), (bytes)) != cudaSuccess) { (p) = nullptr; free_bits_buffers(c); set_error("cudaMalloc failed (bit-stream buffers)"); return P64B_ENOMEM; } \
  if (cudaMemsetAsync((p), 0, (bytes), c->stream) != cudaSuccess) { free_bits_buffers(c); set_error("cudaMemset failed"); return P64B_ECUDA; }
  ALLOCZ(c->d_gob_words, S * ng * VLC_GOB_WORDS * 4);
  ALLOCZ(c->d_gob_bits, S * ng * VLC_PPG * 4);
  ALLOCZ(c->d_carry, S * 4);
  ALLOCZ(c->d_carry_len, S * 4);
  ALLOCZ(c->d_bitpos, S * 8);
  ALLOCZ(c->d_frame_bits, S * 4);
  ALLOCZ(c->d_buffer_offset, S * 8);
  ALLOCZ(c->d_overflows, S * 4);
  ALLOCZ(c->d_pic_hdr, p64b_ctx::NSLOT * 2 * 4);
  c->use_graphs = getenv("P64B_NO_GRAPHS") == nullptr;
  for (int i = 0; i < p64b_ctx::NSLOT; i++) { ALLOCZ(c->d_bits_out[i], c->bits_out_cap); }
  ALLOCZ(c->d_vlc_tables, sizeof(DevVlcTables));
#undef ALLOCZ
  if (cudaMemcpyAsync(c->d_vlc_tables, &t, sizeof t, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess) {       // `t` is on this stack frame
    free_bits_buffers(c); set_error("upload of the VLC tables failed"); return P64B_ECUDA;
  }
  c->bits_budget = S * (c->image_type == P64B_IT_QCIF ? 12u : 40u) * 1024u;   // first frames are intra: generous start
  return 0;
}

static int ensure_host_bits(p64b_ctx* c, int slot, size_t bytes) {
  if (c->h_bits_cap[slot] >= bytes) return 0;
  if (c->h_bits_out[slot]) { CU(cudaStreamSynchronize(c->s_d2h)); CU(cudaFreeHost(c->h_bits_out[slot])); c->h_bits_out[slot] = nullptr; c->h_bits_cap[slot] = 0; }
  const size_t cap = std::min(c->bits_out_cap, bytes + bytes / 2);
  if (cudaHostAlloc((void**)&c->h_bits_out[slot], cap, cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc failed (bit-stream buffer)"); return P64B_ENOMEM; }
  c->h_bits_cap[slot] = cap;
  return 0;
}

// kernel attributes and lazily resolved entry points must exist before a stream capture starts
static int ensure_kernel_attrs(p64b_ctx* c) {
  if (!c->me_attr_done) {
    CU(cudaFuncSetAttribute(me_search_kernel<ME_V_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_FULL));
    CU(cudaFuncSetAttribute(me_search_kernel<ME_V_SURF>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_SURF));
    CU(cudaFuncSetAttribute(me_search_kernel<ME_V_TSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_FULL));
    CU(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, c->device));
    c->me_attr_done = true;
  }
  if (!c->mb_attr_done) {
    CU(cudaFuncSetAttribute(mb_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB4_SMEM));
    c->mb_attr_done = true;
  }
  if (!encode_tiled_fn()) { set_error("cuTensorMapEncodeTiled entry point not available"); return P64B_ECUDA; }
  return 0;
}

// One frame step of every stream under rate control, enqueued on the context's stream (eagerly, or into a capture):
// motion estimation once; then GOB by GOB {quantise..reconstruct with the stream's GQUANT, entropy-code, count the bits,
// choose the next GOB's GQUANT}; macroblocks the overflow test overrode are re-reconstructed as copies at the end
// (macroblocks of a frame do not depend on each other); SwapFS; frame assembly.
static int enqueue_rc_frame(p64b_ctx* c, const p64b_step* st, int slot, VlcArgs a, const VlcFrameArgs& f, const RcArgs& r) {
  int rc;
  const uint8_t* src_dev = c->p_src[slot];
  if (!st->first_frame) {
    if ((rc = launch_me(c, c->d_fs[c->cur], src_dev, (size_t)c->g.frame_bytes, c->S, st->me_mode, st->search_limit, c->d_me))) return rc;
  } else {
    CU(cudaMemsetAsync(c->d_me, 0, (size_t)c->S * c->g.nmb * sizeof(p64b_me), c->stream));
  }
  CU(cudaMemsetAsync(c->d_ovf, 0, (size_t)c->S * c->g.nmb, c->stream));
  rc_frame_begin_kernel<<<(c->S + 255) / 256, 256, 0, c->stream>>>(r, c->S, c->rc_started ? 0 : st->gquant);
  c->rc_started = true;
  c->launches++;
  for (int g = 0; g < c->g.ngob; g++) {
    if ((rc = launch_mb(c, st, src_dev, g, 1, c->d_quant, c->p_mbs[slot], c->p_levels[slot], true))) return rc;
    a.gob_first = g; a.gob_count = 1;
    ProfScope ps(c, 2);
    vlc_gob_kernel<true><<<c->S, VLC_THREADS, 0, c->stream>>>(a);
    c->launches++;
  }
  {
    const int warps = c->S * c->g.nmb, threads = 256;
    overflow_patch_kernel<<<(warps * 32 + threads - 1) / threads, threads, 0, c->stream>>>(
        c->g, c->d_ovf, c->d_fs[c->cur], c->d_fs[c->cur ^ 1], c->d_li[c->cur], c->d_li[c->cur ^ 1], c->S);
  }
  swap_stores(c);
  c->last_src = src_dev;
  {
    ProfScope ps(c, 2);
    vlc_sizes_kernel<<<1, 1024, 0, c->stream>>>(f);
    vlc_frame_kernel<<<c->S, VLC_FRAME_THREADS, 0, c->stream>>>(f);
  }
  c->launches += 3;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int p64b_ctx_set_input_chroma(p64b_ctx* c, int chroma) {
  if (!c) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->submitted || c->frame_src) { set_error("the input format must be configured before the first frame"); return P64B_EINVAL; }
  const int raw = p64b_raw_frame_bytes(c->image_type, chroma);
  if (raw < 0) { set_error("unknown chroma type"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  const size_t wh = (size_t)c->g.W * c->g.H;
  c->chroma = chroma; c->raw_bytes = (size_t)raw;
  c->aux_bytes = chroma == P64B_CHROMA_444ALPHA ? 2 * wh : (size_t)raw - wh;       // the alpha plane is never uploaded
  for (int i = 0; i < p64b_ctx::NSLOT; i++) {
    cudaFree(c->d_aux[i]); c->d_aux[i] = nullptr;
    if (chroma != P64B_CHROMA_420JPEG && c->aux_bytes)
      if (cudaMalloc((void**)&c->d_aux[i], (size_t)c->S * c->aux_bytes + 64) != cudaSuccess) { set_error("cudaMalloc failed (ingest buffers)"); return P64B_ENOMEM; }
  }
  return 0;
}

extern "C" int p64b_ctx_convert_frames(p64b_ctx* c, const uint8_t* raw, uint8_t* out420) {
  if (!c || !raw || !out420) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->frame_src) { set_error("p64b_ctx_convert_frames inside frame_begin/frame_end"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  CU(cudaStreamSynchronize(c->s_h2d));      // staging slot 0 may still be in use by pipelined steps
  CU(cudaStreamSynchronize(c->stream));
  if ((rc = upload_source(c, 0, c->d_src, raw, c->stream)) || (rc = ingest_source(c, 0, c->d_src))) return rc;
  CU(cudaMemcpyAsync(out420, c->d_src, (size_t)c->S * c->g.frame_bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int p64b_ctx_set_rate_control(p64b_ctx* c, const p64b_rate_control* r) {
  if (!c || !r) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->submitted || c->frame_src) { set_error("rate control must be configured before the first frame"); return P64B_EINVAL; }
  if (r->rate < 0 || (r->rate && (r->frame_rate < 1 || r->frame_rate_div < 1 || r->frame_skip < 1 || r->qdfact < 1))) {
    set_error("bad rate-control parameters"); return P64B_EINVAL;
  }
  c->rate = *r;
  return 0;
}

// the download of a bits step: the per-stream table + the stream bytes up to the budget, on the copy-back stream
static int issue_bits_download(p64b_ctx* c, int slot) {
  static const bool header_only = getenv("P64B_BITS_D2H") && !strcmp(getenv("P64B_BITS_D2H"), "header");    // attribution experiments only
  const size_t copy = header_only ? vlc_data_offset(c->S) : std::min(c->bits_out_cap, vlc_data_offset(c->S) + c->bits_budget);
  int rc;
  if ((rc = ensure_host_bits(c, slot, copy))) return rc;
  CU(cudaMemcpyAsync(c->h_bits_out[slot], c->d_bits_out[slot], copy, cudaMemcpyDeviceToHost, c->s_d2h));
  CU(cudaEventRecord(c->ev_d2h[slot], c->s_d2h));
  c->slot_copied[slot] = copy;
  return 0;
}

extern "C" int p64b_ctx_submit_bits(p64b_ctx* c, const p64b_step* st, int temporal_reference, const uint8_t* src, int64_t* ticket) {
  if (!c || !src || !ticket) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->frame_src) { set_error("p64b_ctx_submit_bits inside frame_begin/frame_end"); return P64B_EINVAL; }
  int rc;
  if ((rc = check_step(st)) || (rc = check_quant(st->gquant)) || (rc = use_device(c)) || (rc = ensure_bits_buffers(c))) return rc;
  const int slot = (int)(c->submitted % p64b_ctx::NSLOT);
  if (c->slot_used[slot]) {
    CU(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[slot], 0));    // the slot's source was consumed by its last kernels
    CU(cudaStreamWaitEvent(c->stream, c->ev_d2h[slot], 0));    // its outputs were downloaded
  }
  if ((rc = upload_source(c, slot, c->p_src[slot], src, c->s_h2d))) return rc;
  CU(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
  CU(cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0));
  if (c->pending_d2h >= 0) {                 // the previous step's download starts when this upload has finished
    CU(cudaStreamWaitEvent(c->s_d2h, c->ev_h2d[slot], 0));
    if ((rc = issue_bits_download(c, c->pending_d2h))) return rc;
    c->pending_d2h = -1;
  }
  if ((rc = ingest_source(c, slot, c->p_src[slot]))) return rc;
  // WritePictureHeader, marker.c:103-137: PSC(20) TR(5) PTYPE(6) [PEI=1 PSPARE(8) for NTSC, p64.c:408-423] PEI=0
  VlcFrameArgs f{};
  {
    uint64_t h = (0x10ull << 44) | ((uint64_t)(temporal_reference & 31) << 39) |
                 ((uint64_t)(c->image_type == P64B_IT_QCIF ? 0x00 : 0x04) << 33);
    f.pic_hdr_bits = 32;
    if (c->image_type == P64B_IT_NTSC) { h |= (1ull << 32) | (0x8cull << 24); f.pic_hdr_bits = 41; }
    if (c->rate.rate) {
      f.pic_hdr = c->d_pic_hdr + 2 * slot;
      set_pic_hdr_kernel<<<1, 1, 0, c->stream>>>(c->d_pic_hdr + 2 * slot, (uint32_t)(h >> 32), (uint32_t)h);
      c->launches++;
    } else {
      f.pic_hdr = nullptr; f.pic_hdr_imm[0] = (uint32_t)(h >> 32); f.pic_hdr_imm[1] = (uint32_t)h;
    }
  }
  f.gob_words = c->d_gob_words; f.gob_bits = c->d_gob_bits; f.carry = c->d_carry; f.carry_len = c->d_carry_len;
  f.bitpos = c->d_bitpos; f.out = c->d_bits_out[slot]; f.n_streams = c->S; f.ngob = c->g.ngob; f.gquant = st->gquant;
  VlcArgs a{};
  a.tables = c->d_vlc_tables; a.mbs = c->p_mbs[slot]; a.levels = c->p_levels[slot];
  a.gob_words = c->d_gob_words; a.gob_bits = c->d_gob_bits;
  a.n_streams = c->S; a.ngob = c->g.ngob; a.nmb = c->g.nmb; a.qcif = c->g.qcif; a.gquant = st->gquant;
  if (!c->rate.rate) {
    if ((rc = p64b_ctx_encode_frames_dev(c, st, c->p_src[slot], c->p_mbs[slot], c->p_levels[slot]))) return rc;
    a.gob_first = 0; a.gob_count = c->g.ngob;
    f.ppg = VLC_PPG;
    int grid;
    if ((rc = vlc_seq_grid(c, &a, &grid))) return rc;
    ProfScope ps(c, 2);
    vlc_gob_seq_kernel<<<grid, VLC_SEQ_THREADS, 0, c->stream>>>(a);
    vlc_sizes_kernel<<<1, 1024, 0, c->stream>>>(f);
    vlc_frame_kernel<<<c->S, VLC_FRAME_THREADS, 0, c->stream>>>(f);
    c->launches += 3;
    CU(cudaGetLastError());
  } else {
    // Rate control (-r): motion estimation once per frame; then GOB by GOB {quantise..reconstruct with the stream's GQUANT,
    // entropy-code, count the bits, choose the next GOB's GQUANT} -- all on the device, no host round trip.  Macroblocks the
    // overflow test overrode are re-reconstructed as copies at the end (macroblocks of a frame do not depend on each other).
    RcArgs r{};
    r.rate = c->rate.rate; r.frame_skip = c->rate.frame_skip; r.frame_rate = c->rate.frame_rate;
    r.frame_rate_div = c->rate.frame_rate_div; r.qdfact = c->rate.qdfact; r.qoffs = c->rate.qoffs;
    r.denom = c->g.ngob * 33 * c->rate.frame_rate / c->rate.frame_rate_div;
    r.first_frame = st->first_frame; r.pic_hdr_bits = f.pic_hdr_bits;
    r.bitpos = c->d_bitpos; r.frame_bits = c->d_frame_bits; r.buffer_offset = c->d_buffer_offset;
    r.quant = c->d_quant; r.ovf = c->d_ovf; r.overflows = c->d_overflows;
    a.rc = r; f.rc = r;
    if ((rc = ensure_kernel_attrs(c))) return rc;
    const bool graphable = c->use_graphs && !st->first_frame && c->rc_started && !c->prof;
    if (!graphable) {
      if ((rc = enqueue_rc_frame(c, st, slot, a, f, r))) return rc;
    } else {
      const int parity = (int)(c->me_launches & 1);
      p64b_ctx::RcGraph* g = nullptr;
      for (auto& e : c->rc_graphs)
        if (e.slot == slot && e.cur == c->cur && e.parity == parity && e.me_mode == st->me_mode && e.search_limit == st->search_limit &&
            e.force_intra == st->force_intra) { g = &e; break; }
      if (g) {                                       // replay: the same launches with the same arguments
        CU(cudaGraphLaunch(g->exec, c->stream));
        c->me_launches++; c->launches += g->kernels;
        swap_stores(c);
        c->last_src = c->p_src[slot];
      } else {
        const int64_t l0 = c->launches, ml0 = c->me_launches;
        const int cur0 = c->cur;
        const uint8_t* last0 = c->last_src;
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_rc_frame(c, st, slot, a, f, r);
        const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc || ce != cudaSuccess || !graph) {     // nothing was enqueued: the context's frame-store / queue state is as before
          if (graph) cudaGraphDestroy(graph);
          c->cur = cur0; c->me_launches = ml0; c->launches = l0; c->last_src = last0;
          if (!rc) { set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); rc = P64B_ECUDA; }
          return rc;
        }
        p64b_ctx::RcGraph e{slot, c->cur ^ 1, parity, st->me_mode, st->search_limit, st->force_intra, nullptr, (int)(c->launches - l0)};
        const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
          c->cur = cur0; c->me_launches = ml0; c->launches = l0; c->last_src = last0;
          set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); return P64B_ECUDA;
        }
        c->rc_graphs.push_back(e);                   // (enqueue_rc_frame already flipped c->cur: the key is the state BEFORE the step)
        CU(cudaGraphLaunch(e.exec, c->stream));
      }
    }
  }
  CU(cudaEventRecord(c->ev_comp[slot], c->stream));
  CU(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[slot], 0));
  // The download of this step is not started now but when the NEXT step's upload has finished (or at p64b_ctx_wait_bits, if no
  // further step was submitted): a download that falls into the last part of an upload slows that upload down (measured on
  // one B200: 340 k -> 357 k frames/s end to end, 0.94 -> 0.985 of the upload bound), one that starts
  // together with an upload does not.  P64B_D2H_IMMEDIATE=1 restores the immediate download (attribution only).
  static const bool defer = getenv("P64B_D2H_IMMEDIATE") == nullptr;
  if (defer) c->pending_d2h = slot;
  else if ((rc = issue_bits_download(c, slot))) return rc;
  c->slot_used[slot] = true;
  c->slot_bits[slot] = true;
  *ticket = c->submitted++;
  return 0;
}

extern "C" int p64b_ctx_encode_bits_dev(p64b_ctx* c, const p64b_step* st, int temporal_reference, const uint8_t* src_dev) {
  if (!c || !src_dev) { set_error("NULL argument"); return P64B_EINVAL; }
  if (c->frame_src) { set_error("p64b_ctx_encode_bits_dev inside frame_begin/frame_end"); return P64B_EINVAL; }
  if (c->rate.rate) { set_error("p64b_ctx_encode_bits_dev: fixed quantiser only"); return P64B_EINVAL; }
  int rc;
  if ((rc = check_step(st)) || (rc = check_quant(st->gquant)) || (rc = use_device(c)) || (rc = ensure_bits_buffers(c))) return rc;
  VlcFrameArgs f{};
  uint64_t h = (0x10ull << 44) | ((uint64_t)(temporal_reference & 31) << 39) | ((uint64_t)(c->image_type == P64B_IT_QCIF ? 0x00 : 0x04) << 33);
  f.pic_hdr_bits = 32;
  if (c->image_type == P64B_IT_NTSC) { h |= (1ull << 32) | (0x8cull << 24); f.pic_hdr_bits = 41; }
  f.pic_hdr = nullptr; f.pic_hdr_imm[0] = (uint32_t)(h >> 32); f.pic_hdr_imm[1] = (uint32_t)h;
  f.gob_words = c->d_gob_words; f.gob_bits = c->d_gob_bits; f.carry = c->d_carry; f.carry_len = c->d_carry_len;
  f.bitpos = c->d_bitpos; f.out = c->d_bits_out[0]; f.n_streams = c->S; f.ngob = c->g.ngob; f.gquant = st->gquant;
  VlcArgs a{};
  a.tables = c->d_vlc_tables; a.mbs = c->p_mbs[0]; a.levels = c->p_levels[0];
  a.gob_words = c->d_gob_words; a.gob_bits = c->d_gob_bits;
  a.n_streams = c->S; a.ngob = c->g.ngob; a.nmb = c->g.nmb; a.qcif = c->g.qcif; a.gquant = st->gquant;
  a.gob_first = 0; a.gob_count = c->g.ngob;
  if ((rc = p64b_ctx_encode_frames_dev(c, st, src_dev, c->p_mbs[0], c->p_levels[0]))) return rc;
  f.ppg = VLC_PPG;
  int grid;
  if ((rc = vlc_seq_grid(c, &a, &grid))) return rc;
  ProfScope ps(c, 2);
  vlc_gob_seq_kernel<<<grid, VLC_SEQ_THREADS, 0, c->stream>>>(a);
  vlc_sizes_kernel<<<1, 1024, 0, c->stream>>>(f);
  vlc_frame_kernel<<<c->S, VLC_FRAME_THREADS, 0, c->stream>>>(f);
  c->launches += 3;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int p64b_ctx_wait_bits(p64b_ctx* c, int64_t ticket, p64b_bits_out* out) {
  if (!c || !out || ticket < 0 || ticket >= c->submitted || ticket + p64b_ctx::NSLOT < c->submitted) { set_error("bad ticket"); return P64B_EINVAL; }
  int rc;
  if ((rc = use_device(c))) return rc;
  const int slot = (int)(ticket % p64b_ctx::NSLOT);
  if (!c->slot_bits[slot]) { set_error("ticket was not issued by p64b_ctx_submit_bits"); return P64B_EINVAL; }
  if (c->pending_d2h == slot) {              // nobody submitted a further step: download now
    if ((rc = issue_bits_download(c, slot))) return rc;
    c->pending_d2h = -1;
  }
  if (!c->h_bits_out[slot]) { set_error("ticket was not issued by p64b_ctx_submit_bits"); return P64B_EINVAL; }
  CU(cudaEventSynchronize(c->ev_d2h[slot]));
  const size_t doff = vlc_data_offset(c->S);
  size_t total = reinterpret_cast<const uint32_t*>(c->h_bits_out[slot])[c->S];
  static const bool header_only = getenv("P64B_BITS_D2H") && !strcmp(getenv("P64B_BITS_D2H"), "header");
  if (doff + total > c->slot_copied[slot] && !header_only) {
    // the frame was larger than the download budget: fetch the rest (rare; the budget follows the frame sizes)
    const size_t have = c->slot_copied[slot];
    if (c->h_bits_cap[slot] < doff + total) {
      uint8_t* nb = nullptr;
      const size_t cap = std::min(c->bits_out_cap, doff + total + total / 2);
      if (cudaHostAlloc((void**)&nb, cap, cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc failed (bit-stream buffer)"); return P64B_ENOMEM; }
      memcpy(nb, c->h_bits_out[slot], have);
      CU(cudaFreeHost(c->h_bits_out[slot]));
      c->h_bits_out[slot] = nb; c->h_bits_cap[slot] = cap;
    }
    // (on the copy-back stream this copy would queue behind the downloads of the steps already submitted, i.e. behind their
    // kernels: the context's upload stream is idle-waiting far less often and keeps the pipeline depth)
    CU(cudaMemcpyAsync(c->h_bits_out[slot] + have, c->d_bits_out[slot] + have, doff + total - have, cudaMemcpyDeviceToHost, c->s_fix));
    CU(cudaStreamSynchronize(c->s_fix));
    c->slot_copied[slot] = doff + total;
    c->second_copies++;
  }
  // next step's first copy: the largest of the last four frames + 1/32 + 16 KB (the size of a whole batch's frame moves slowly;
  // a frame that outgrows the budget is completed by the second copy above, which stalls the host for a copy's latency).  Downloads share the host link with the uploads, and on a box
  // whose GPUs share uplinks every downloaded byte costs upload time: no generous slack.
  c->bits_recent[ticket & 3] = total;
  const size_t recent = std::max(std::max(c->bits_recent[0], c->bits_recent[1]), std::max(c->bits_recent[2], c->bits_recent[3]));
  c->bits_budget = std::max<size_t>(recent + recent / 32 + 16384, 65536);
  const uint8_t* b = c->h_bits_out[slot];
  const size_t S = (size_t)c->S;
  out->offset = reinterpret_cast<const uint32_t*>(b);
  out->nbytes = out->offset + (S + 1);
  out->carry = out->nbytes + S;
  out->carry_len = out->carry + S;
  out->bit_position = reinterpret_cast<const uint64_t*>(b + vlc_bitpos_offset(c->S));
  out->gquant = reinterpret_cast<const uint32_t*>(b + vlc_bitpos_offset(c->S) + S * 8);
  out->overflows = out->gquant + S;
  out->data = b + doff;
  out->total_bytes = total;
  out->downloaded_bytes = c->slot_copied[slot];
  return 0;
}

// Host-to-device copy rate of `bytes` from (pinned) host memory, `reps` copies back to back on one stream: what the PCIe
// link gives the upload-bound end-to-end path.
int p64b_measure_h2d(int device, const void* host, size_t bytes, int reps, double* gb_per_s) {
  if (!host || !gb_per_s || !bytes || reps < 1) return P64B_EINVAL;
  CU(cudaSetDevice(device));
  void* d = nullptr;
  CU(cudaMalloc(&d, bytes));
  cudaStream_t st;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CU(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; i++) CU(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st); cudaFree(d);
  *gb_per_s = (double)bytes * reps / (best * 1e-3) / 1e9;
  return 0;
}

int p64b_measure_link(int device, const void* const* up, int up_sets, size_t up_bytes, void* down, size_t down_bytes, int reps,
                      int mode, double* up_gb_per_s, double* down_gb_per_s) {
  const bool do_up = mode & 1, do_down = mode & 2;
  if (reps < 1 || (!do_up && !do_down) || (do_up && (!up || up_sets < 1 || !up_bytes)) || (do_down && (!down || !down_bytes))) return P64B_EINVAL;
  CU(cudaSetDevice(device));
  void *d_up = nullptr, *d_down = nullptr;
  if (do_up) CU(cudaMalloc(&d_up, up_bytes));
  if (do_down) { CU(cudaMalloc(&d_down, down_bytes)); CU(cudaMemset(d_down, 1, down_bytes)); }
  cudaStream_t su, sd;
  CU(cudaStreamCreateWithFlags(&su, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
  cudaEvent_t u0, u1, d0, d1;
  CU(cudaEventCreate(&u0)); CU(cudaEventCreate(&u1)); CU(cudaEventCreate(&d0)); CU(cudaEventCreate(&d1));
  CU(cudaDeviceSynchronize());
  if (do_up) CU(cudaEventRecord(u0, su));
  if (do_down) CU(cudaEventRecord(d0, sd));
  for (int i = 0; i < reps; i++) {
    if (do_up) CU(cudaMemcpyAsync(d_up, up[i % up_sets], up_bytes, cudaMemcpyHostToDevice, su));
    if (do_down) CU(cudaMemcpyAsync(down, d_down, down_bytes, cudaMemcpyDeviceToHost, sd));
  }
  if (do_up) CU(cudaEventRecord(u1, su));
  if (do_down) CU(cudaEventRecord(d1, sd));
  CU(cudaStreamSynchronize(su)); CU(cudaStreamSynchronize(sd));
  float ms = 0;
  if (up_gb_per_s) { *up_gb_per_s = 0; if (do_up) { CU(cudaEventElapsedTime(&ms, u0, u1)); *up_gb_per_s = (double)up_bytes * reps / (ms * 1e-3) / 1e9; } }
  if (down_gb_per_s) { *down_gb_per_s = 0; if (do_down) { CU(cudaEventElapsedTime(&ms, d0, d1)); *down_gb_per_s = (double)down_bytes * reps / (ms * 1e-3) / 1e9; } }
  cudaEventDestroy(u0); cudaEventDestroy(u1); cudaEventDestroy(d0); cudaEventDestroy(d1);
  cudaStreamDestroy(su); cudaStreamDestroy(sd); cudaFree(d_up); cudaFree(d_down);
  return 0;
}

int p64b_debug_oob(int device, uint32_t* violations, uint32_t* last_line) {
  unsigned int v[2] = {0, 0};
  if (cudaSetDevice(device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpyFromSymbol(v, g_oob, sizeof v) != cudaSuccess) { set_error("p64b_debug_oob: device not usable"); return P64B_ECUDA; }
  if (violations) *violations = v[0];
  if (last_line) *last_line = v[1];
#ifdef P64B_BOUNDS_CHECK
  return 1;
#else
  return 0;
#endif
}

int p64b_probe_links(const int32_t* devices, int n_devices, double* gb_per_s) {
  if (!devices || !gb_per_s || n_devices < 1) return P64B_EINVAL;
  const size_t bytes = 32u << 20;
  std::vector<void*> host(n_devices, nullptr);
  std::vector<int> rc(n_devices, 0);
  for (int k = 0; k < n_devices; k++) {
    if (cudaSetDevice(devices[k]) != cudaSuccess || !(host[k] = p64b_host_alloc(bytes))) {
      for (void* h : host) p64b_host_free(h);
      set_error("p64b_probe_links: bad device or no pinned memory");
      return P64B_ECUDA;
    }
    memset(host[k], k + 1, bytes);
  }
  std::vector<std::thread> th;
  for (int k = 0; k < n_devices; k++)
    th.emplace_back([&, k] {
      const void* up[1] = {host[k]};
      double d = 0;
      p64b_measure_link(devices[k], up, 1, bytes, nullptr, 0, 2, 1, &gb_per_s[k], &d);                     // warm-up (context, first touch)
      rc[k] = p64b_measure_link(devices[k], up, 1, bytes, nullptr, 0, 8, 1, &gb_per_s[k], &d);
    });
  for (auto& t : th) t.join();
  for (void* h : host) p64b_host_free(h);
  for (int k = 0; k < n_devices; k++) if (rc[k]) { set_error("p64b_probe_links: copy failed"); return rc[k]; }
  return 0;
}

int p64b_measure_sad_peak(int device, double* ops_per_s, double* sm_clock_mhz) {
  if (!ops_per_s) return P64B_EINVAL;
  CU(cudaSetDevice(device));
  int sms = 0, khz = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CU(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
  uint32_t* d = nullptr;
  const int blocks = sms * 8, iters = 4000;
  CU(cudaMalloc((void**)&d, sizeof(uint32_t) * blocks * 256));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  sad_peak_kernel<<<blocks, 256>>>(d, 200, 1u);
  CU(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CU(cudaEventRecord(e0));
    sad_peak_kernel<<<blocks, 256>>>(d, iters, (uint32_t)(r + 2));
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *ops_per_s = (double)blocks * 256.0 * iters * 64.0 / (best * 1e-3);
  if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
  return 0;
}

}  // extern "C"
