// Sequence driver (p64b_enc_*): the reference's p64EncodeSequence / p64EncodeFrame / p64EncodeGOB host
// control (p64.c:524-786) for a batch of independent streams.  Everything data-parallel is delegated to
// the device context (p64b_ctx_*); what stays here is exactly what the north star keeps sequential:
// headers + VLC (bits.cpp) and rate control (ExecuteQuantization p64.c:458-481, BufferContents p64.c:233-237,
// the per-MB overflow test p64.c:776-783).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/p64_b200.h"

namespace p64b { void set_error(const std::string& s); }

namespace {

struct StreamState {
  p64b_bits* bits = nullptr;
  int gquant = 8;            // GQuant (p64.c:89)
  int64_t buffer_offset = 0; // BufferOffset (p64.c:140)
  int64_t total_bits = 0, first_frame_bits = 0, overflows = 0;
  int64_t last_bits = 0, buffer_contents_at_end = 0;   // LastBits; BufferContents() when PrintFrameStatistics runs
  // device-side entropy coding: the stream's bytes as the device returned them, and its pending bits (< 8)
  std::vector<uint8_t> dev_bytes;
  uint32_t carry = 0, carry_len = 0;
};

// One host worker thread per device partition of a multi-device encoder: runs the partition's own p64b_enc (its own
// p64b_ctx, pipeline and per-stream state) on request.  No data crosses partitions.
struct Worker {
  p64b_enc* kid = nullptr;
  int first_stream = 0;
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  enum Task { NONE, ENCODE, FINISH, QUIT } task = NONE;
  const uint8_t* src = nullptr;
  bool busy = false;
  int rc = 0;
  std::string err;
};

}  // namespace

// A few host threads that stay alive for the encoder's lifetime: the per-frame host work (appending every stream's bytes,
// the host VLC, the staging copy) is a parallel loop of well under a millisecond, and starting threads for each one costs
// more than the loop (300 frames x 256 streams: 0.9 -> 2.7 ms per step from run to run with per-call threads).
struct Pool {
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread> th;
  const std::function<void(int)>* fn = nullptr;
  int n = 0;
  std::atomic<int> next{0};
  int running = 0;                 // workers that have not yet left the current loop
  uint64_t gen = 0;
  bool quit = false;
  explicit Pool(int workers) {
    for (int i = 0; i < workers; i++) th.emplace_back([this] { loop(); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(m); quit = true; }
    cv_work.notify_all();
    for (auto& t : th) t.join();
  }
  void loop() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv_work.wait(lk, [&] { return quit || gen != seen; });
      if (quit) return;
      seen = gen;
      const std::function<void(int)>* f = fn;
      const int cnt = n;
      lk.unlock();
      for (int i; (i = next.fetch_add(1)) < cnt;) (*f)(i);
      lk.lock();
      if (--running == 0) cv_done.notify_one();
    }
  }
  // f(0) .. f(cnt-1), each once, on the workers and the caller; returns when all are done
  void run(int cnt, const std::function<void(int)>& f) {
    { std::lock_guard<std::mutex> lk(m); fn = &f; n = cnt; next.store(0); running = (int)th.size(); gen++; }
    cv_work.notify_all();
    for (int i; (i = next.fetch_add(1)) < cnt;) f(i);
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return running == 0; });
  }
};

struct p64b_enc {
  p64b_enc_params p{};
  // multi-device parent: the partitions; everything per stream lives in the kids
  std::vector<std::unique_ptr<Worker>> workers;
  bool external_stage = false;   // kid of a multi-device encoder: `src` is the parent's pinned staging, used in place
  p64b_ctx* ctx = nullptr;
  int S = 0, ngob = 0, nmb = 0, frame_bytes = 0;
  int src_bytes = 0;             // bytes per stream of one input frame: frame_bytes, or the raw Y4M payload (input_chroma)
  int current_frame = 0;         // CurrentFrame (p64.c:121)
  int frames_done = 0;
  int qdfact = 1, qoffs = 1;     // QDFact, QOffs (p64.c:141-142)
  bool finished = false;
  std::vector<StreamState> st;
  p64b_mb* h_mbs = nullptr;      // pinned [S][nmb]
  int8_t* h_levels = nullptr;    // pinned [S][nmb][384]
  uint8_t* h_src = nullptr;      // pinned staging [S][src_bytes] of the NEXT frame (what p64b_enc_staging returns)
  // device path: up to three frames in flight (p64b_ctx_submit_bits), harvested in order; four staging buffers so that
  // the one handed out for the next frame is never one a pending upload still reads
  static constexpr int DEPTH = 3, NSTAGE = 4;
  uint8_t* h_ring[NSTAGE] = {};
  int64_t tickets[DEPTH] = {};
  int harvested = 0;             // frames whose bytes have been appended to the streams
  std::vector<uint8_t> quant, overflow;
  int threads = 1;
  std::unique_ptr<Pool> pool;    // created at the first parallel loop
  bool device_vlc = false;       // headers + VLC (+ rate control) run on the device (p64b_ctx_submit_bits)
};

namespace {

Pool* pool_of(p64b_enc* e) {
  // the device path's host work is memory-bound (a few MB per frame): eight threads; the host VLC takes what it is given
  if (!e->pool) e->pool.reset(new Pool(std::max(1, e->device_vlc ? std::min(e->threads, 8) : e->threads) - 1));
  return e->pool.get();
}

template <class F>
void parallel_streams(p64b_enc* e, F f) {
  const int T = std::min(e->threads, e->S);
  if (T <= 1) { for (int s = 0; s < e->S; s++) f(s); return; }
  const int per = e->device_vlc ? 8 : 1;               // streams per work item
  const std::function<void(int)> job = [&](int i) { for (int s = i * per; s < std::min(e->S, (i + 1) * per); s++) f(s); };
  pool_of(e)->run((e->S + per - 1) / per, job);
}

// the caller's frame set into pinned staging: 39 MB per step at 256 CIF streams -- one core copies that in 3 ms, which would
// be the slowest stage of the pipeline; a few threads bring it under the upload time
void copy_frames(p64b_enc* e, uint8_t* dst, const uint8_t* src, size_t bytes) {
  const int T = (int)std::min<size_t>(std::min(e->threads, 8), bytes >> 20);
  if (T <= 1) { memcpy(dst, src, bytes); return; }
  const size_t chunk = (size_t)1 << 20;
  const std::function<void(int)> job = [&](int i) { const size_t o = (size_t)i * chunk; memcpy(dst + o, src + o, std::min(chunk, bytes - o)); };
  pool_of(e)->run((int)((bytes + chunk - 1) / chunk), job);
}

// BufferContents(), p64.c:233-237, with CurrentGOB=g, CurrentMDU=m. int arithmetic as in the reference.
inline int64_t buffer_contents(const p64b_enc* e, const StreamState& s, int g, int m) {
  const int denom = e->ngob * 33 * e->p.frame_rate / e->p.frame_rate_div;
  const int num = (int)((int64_t)(g * 33 + m) * e->p.rate * e->p.frame_skip);
  return p64b_bits_tell(s.bits) + s.buffer_offset - (denom ? num / denom : 0);
}
inline int buffer_size(const p64b_enc* e) { return e->p.rate / 4; }   // BufferSize(), p64.c:237

// ExecuteQuantization(), p64.c:458-481 (no interpreter): GQuant from the buffer fullness.
inline int execute_quantization(const p64b_enc* e, const StreamState& s, int g, int m) {
  int cur = (int)buffer_contents(e, s, g, m);
  int q = cur / e->qdfact + e->qoffs;
  return std::min(std::max(q, 1), 31);
}

// device path: collect frame k's bytes (submitted earlier) into the streams
int harvest(p64b_enc* e, int k) {
  p64b_bits_out o{};
  int rc = p64b_ctx_wait_bits(e->ctx, e->tickets[k % p64b_enc::DEPTH], &o);
  if (rc) return rc;
  parallel_streams(e, [&](int s) {
    StreamState& ss = e->st[s];
    ss.dev_bytes.insert(ss.dev_bytes.end(), o.data + o.offset[s], o.data + o.offset[s] + o.nbytes[s]);
    ss.carry = o.carry[s]; ss.carry_len = o.carry_len[s];
    const int64_t before = ss.total_bits;
    ss.total_bits = (int64_t)o.bit_position[s];
    ss.last_bits = ss.total_bits - before;
    if (k == 0) ss.first_frame_bits = ss.total_bits;
    ss.gquant = (int)o.gquant[s]; ss.overflows = (int64_t)o.overflows[s];
  });
  e->harvested = k + 1;
  return 0;
}

void worker_main(Worker* w) {
  for (;;) {
    Worker::Task t;
    const uint8_t* src;
    {
      std::unique_lock<std::mutex> lk(w->m);
      w->cv.wait(lk, [&] { return w->busy; });
      t = w->task; src = w->src;
    }
    int rc = 0;
    if (t == Worker::ENCODE) rc = p64b_enc_encode(w->kid, src);
    else if (t == Worker::FINISH) rc = p64b_enc_finish(w->kid);
    {
      std::lock_guard<std::mutex> lk(w->m);
      w->rc = rc;
      if (rc) w->err = p64b_last_error();          // the error text is per thread: hand it to the caller's
      w->busy = false;
    }
    w->cv.notify_all();
    if (t == Worker::QUIT) return;
  }
}

// run `task` on every partition at once, wait for all; first failure wins
int run_all(p64b_enc* e, Worker::Task task, const uint8_t* src, size_t src_bytes) {
  for (auto& w : e->workers) {
    std::lock_guard<std::mutex> lk(w->m);
    w->task = task; w->src = src ? src + (size_t)w->first_stream * src_bytes : nullptr; w->busy = true;
    w->cv.notify_all();
  }
  int rc = 0;
  for (auto& w : e->workers) {
    std::unique_lock<std::mutex> lk(w->m);
    w->cv.wait(lk, [&] { return !w->busy; });
    if (w->rc && !rc) { rc = w->rc; p64b::set_error(w->err); }
  }
  return rc;
}

// stream of a multi-device encoder -> (partition's encoder, local stream); single-device: itself
const p64b_enc* route(const p64b_enc* e, int* stream) {
  if (e->workers.empty()) return e;
  for (auto it = e->workers.rbegin(); it != e->workers.rend(); ++it)
    if (*stream >= (*it)->first_stream) { *stream -= (*it)->first_stream; return (*it)->kid; }
  return e->workers.front()->kid;
}

}  // namespace

extern "C" {

void p64b_enc_default_params(p64b_enc_params* p) {
  memset(p, 0, sizeof(*p));
  p->image_type = P64B_IT_NTSC;       // the reference's default (p64.c:114)
  p->n_streams = 1;
  p->frame_rate = 30000; p->frame_rate_div = 1001;   // p64.c:128-129
  p->frame_skip = 1;
  p->me_mode = P64B_ME_TSS;           // the stock search (me.c:352)
  p->search_limit = 15;               // me.c:62
}

int p64b_enc_create(p64b_enc** out, const p64b_enc_params* p) {
  if (!out || !p) { p64b::set_error("NULL argument"); return P64B_EINVAL; }
  if (p->n_streams < 1 || p->frame_rate < 1 || p->frame_rate_div < 1 || p->frame_skip < 1 ||
      p->initial_quant < 0 || p->initial_quant > 31 || p->rate < 0 || (p->rate > 0 && p->rate < 320)) {
    p64b::set_error("bad encoder parameters"); return P64B_EINVAL;
  }
  if (p->n_devices < 0 || p->n_devices > P64B_MAX_DEVICES) { p64b::set_error("bad device list"); return P64B_EINVAL; }
  if (p->n_devices > 0) {
    // ---- multi-device: contiguous stream blocks (sizes differ by at most one), one encoder + worker thread per block
    p64b_enc* e = new p64b_enc();
    e->p = *p;
    e->S = p->n_streams; e->ngob = p64b_num_gob(p->image_type); e->nmb = p64b_num_mb(p->image_type);
    e->frame_bytes = p64b_frame_bytes(p->image_type);
    e->src_bytes = p->input_chroma != P64B_CHROMA_420JPEG ? p64b_raw_frame_bytes(p->image_type, p->input_chroma) : e->frame_bytes;
    if (e->ngob < 0 || e->src_bytes < 0) { delete e; p64b::set_error("unknown image or chroma type"); return P64B_EINVAL; }
    e->current_frame = p->start_frame;
    e->device_vlc = !p->host_vlc;
    e->threads = p->vlc_threads > 0 ? p->vlc_threads : std::max(1, std::min((int)std::thread::hardware_concurrency(), 64));
    const int nd = p->n_devices, base = p->n_streams / nd, extra = p->n_streams % nd;
    int first = 0, rc = 0;
    std::vector<int> share(nd);
    for (int k = 0; k < nd; k++) share[k] = base + (k < extra ? 1 : 0);
    if (p->balance_links && nd > 1 && p->n_streams >= 2 * nd) {
      // blocks in proportion to what each device's host link delivers while all upload (largest remainder, at least one stream)
      std::vector<double> gbs(nd, 0.0);
      if ((rc = p64b_probe_links(p->devices, nd, gbs.data()))) { delete e; return rc; }
      double sum = 0;
      for (double g : gbs) sum += g;
      int given = 0;
      std::vector<std::pair<double, int>> frac;
      for (int k = 0; k < nd; k++) {
        const double raw = sum > 0 ? gbs[k] / sum * p->n_streams : (double)share[k];
        share[k] = std::max(1, (int)raw);
        given += share[k];
        frac.emplace_back(raw - (int)raw, k);
      }
      std::sort(frac.rbegin(), frac.rend());
      for (int i = 0; given < p->n_streams; i++, given++) share[frac[i % nd].second]++;
      for (int i = 0; given > p->n_streams; i++) { int k = frac[(nd - 1 - i % nd)].second; if (share[k] > 1) { share[k]--; given--; } }
    }
    for (int k = 0; k < nd && !rc; k++) {
      const int n = share[k];
      if (n == 0) continue;                          // fewer streams than devices
      p64b_enc_params kp = *p;
      kp.n_devices = 0; kp.device = p->devices[k]; kp.n_streams = n;
      std::unique_ptr<Worker> w(new Worker());
      w->first_stream = first;
      if ((rc = p64b_enc_create(&w->kid, &kp))) break;
      w->kid->external_stage = true;                  // the partition reads the parent's staging in place: its own ring is not needed
      for (auto*& b : w->kid->h_ring) { p64b_host_free(b); b = nullptr; }
      w->kid->h_src = nullptr;
      e->workers.push_back(std::move(w));
      first += n;
    }
    for (int i = 0; i < p64b_enc::NSTAGE && !rc; i++)
      if (!(e->h_ring[i] = (uint8_t*)p64b_host_alloc((size_t)e->S * e->src_bytes))) rc = P64B_ENOMEM;
    for (auto& w : e->workers) w->th = std::thread(worker_main, w.get());
    if (rc) { p64b_enc_destroy(e); return rc; }
    e->h_src = e->h_ring[0];
    *out = e;
    return 0;
  }
  p64b_enc* e = new p64b_enc();
  e->p = *p;
  int rc = p64b_ctx_create(&e->ctx, p->device, p->image_type, p->n_streams);
  if (rc) { delete e; return rc; }
  e->S = p->n_streams; e->ngob = p64b_num_gob(p->image_type); e->nmb = p64b_num_mb(p->image_type);
  e->frame_bytes = p64b_frame_bytes(p->image_type);
  e->src_bytes = e->frame_bytes;
  if (p->input_chroma != P64B_CHROMA_420JPEG) {      // ingest: chroma conversion on the device (y4m_input.c:195-545)
    e->src_bytes = p64b_raw_frame_bytes(p->image_type, p->input_chroma);
    if (e->src_bytes < 0 || (rc = p64b_ctx_set_input_chroma(e->ctx, p->input_chroma))) {
      if (e->src_bytes < 0) p64b::set_error("unknown input chroma type");
      p64b_ctx_destroy(e->ctx); delete e; return rc ? rc : P64B_EINVAL;
    }
  }
  e->current_frame = p->start_frame;
  int iq = p->initial_quant;                       // p64.c:574-590
  if (p->rate) {
    e->qdfact = p->rate / 320; e->qoffs = 1;
    if (!iq) iq = std::min(std::max(10000000 / p->rate, 1), 31);
  }
  if (!iq) iq = 8;                                 // DEFAULT_QUANTIZATION
  e->st.resize(e->S);
  for (auto& s : e->st) { s.bits = p64b_bits_create(p->image_type); s.gquant = iq; }
  e->h_mbs = (p64b_mb*)p64b_host_alloc((size_t)e->S * e->nmb * sizeof(p64b_mb));
  e->h_levels = (int8_t*)p64b_host_alloc((size_t)e->S * e->nmb * P64B_LEVELS_PER_MB);
  e->device_vlc = !p->host_vlc;
  for (int i = 0; i < (e->device_vlc ? p64b_enc::NSTAGE : 1); i++)
    if (!(e->h_ring[i] = (uint8_t*)p64b_host_alloc((size_t)e->S * e->src_bytes))) { p64b_enc_destroy(e); return P64B_ENOMEM; }
  e->h_src = e->h_ring[0];
  if (!e->h_mbs || !e->h_levels) { p64b_enc_destroy(e); return P64B_ENOMEM; }
  e->quant.assign(e->S, (uint8_t)iq);
  e->overflow.assign((size_t)e->S * e->nmb, 0);
  int hw = (int)std::thread::hardware_concurrency();
  e->threads = p->vlc_threads > 0 ? p->vlc_threads : std::max(1, std::min(hw, 64));
  if (e->device_vlc && p->rate) {                  // the buffer model runs on the device, per stream
    p64b_rate_control r{};
    r.rate = p->rate; r.frame_rate = p->frame_rate; r.frame_rate_div = p->frame_rate_div; r.frame_skip = p->frame_skip;
    r.qdfact = e->qdfact; r.qoffs = e->qoffs;
    if ((rc = p64b_ctx_set_rate_control(e->ctx, &r))) { p64b_enc_destroy(e); return rc; }
  }
  *out = e;
  return 0;
}

int p64b_debug_pool_selftest(int workers, int items, int rounds) {
  if (workers < 0 || workers > 256 || items < 0 || rounds < 0) { p64b::set_error("bad argument"); return P64B_EINVAL; }
  Pool pool(workers);
  std::vector<std::atomic<int>> hits(items);
  for (auto& h : hits) h.store(0);
  for (int r = 0; r < rounds; r++) {
    const std::function<void(int)> job = [&](int i) { hits[i].fetch_add(1); };
    pool.run(items, job);
    for (int i = 0; i < items; i++)
      if (hits[i].load() != r + 1) { p64b::set_error("thread pool: an item did not run exactly once"); return P64B_EINVAL; }
  }
  return 0;
}

void p64b_enc_destroy(p64b_enc* e) {
  if (!e) return;
  for (auto& w : e->workers) {
    if (w->th.joinable()) {
      { std::lock_guard<std::mutex> lk(w->m); w->task = Worker::QUIT; w->busy = true; }
      w->cv.notify_all();
      w->th.join();
    }
    p64b_enc_destroy(w->kid);                      // (drains its pipeline before the shared staging below goes away)
  }
  e->workers.clear();
  for (auto& s : e->st) p64b_bits_destroy(s.bits);
  p64b_host_free(e->h_mbs); p64b_host_free(e->h_levels);
  p64b_ctx_destroy(e->ctx);                        // (drains whatever is still in flight before the staging goes away)
  e->ctx = nullptr;
  for (auto* b : e->h_ring) p64b_host_free(b);
  delete e;
}

int p64b_enc_encode(p64b_enc* e, const uint8_t* src) {
  if (!e || !src) { p64b::set_error("NULL argument"); return P64B_EINVAL; }
  if (e->finished) { p64b::set_error("encoder already finished"); return P64B_EINVAL; }
  if (!e->workers.empty()) {
    // multi-device: the frame goes into the parent's pinned staging ring (four deep, like a single encoder's: the buffer
    // handed out next was last read by frame n-3, which every partition has collected when this call returns), and
    // every partition encodes its block of streams from there, in place, on its own thread and device.
    const int n = e->frames_done;
    uint8_t* stage = e->h_ring[n % p64b_enc::NSTAGE];
    if (src != stage) copy_frames(e, stage, src, (size_t)e->S * e->src_bytes);
    const int rc = run_all(e, Worker::ENCODE, stage, (size_t)e->src_bytes);
    if (rc) return rc;
    e->h_src = e->h_ring[(n + 1) % p64b_enc::NSTAGE];
    e->frames_done++;
    e->current_frame += e->p.frame_skip;
    return 0;
  }
  const bool first = e->frames_done == 0;          // CurrentFrame == StartFrame
  p64b_step step{};
  step.first_frame = first; step.me_mode = e->p.me_mode; step.search_limit = e->p.search_limit;
  step.force_intra = e->p.force_intra;
  const int tr = e->current_frame % 32;            // p64.c:637
  if (!e->device_vlc) for (auto& ss : e->st) p64b_bits_counters_reset(ss.bits);      // p64.c:640-649
  int rc;
  if (e->device_vlc) {
    // device-side entropy coding (and rate control, if any): a device step returns every stream's next whole bytes.  Frames
    // are pipelined: this call enqueues frame n and collects frame n-3, so uploads, kernels and downloads of neighbouring
    // frames overlap; p64b_enc_finish() collects the rest.
    const int n = e->frames_done;
    if (n >= p64b_enc::DEPTH && (rc = harvest(e, n - p64b_enc::DEPTH))) return rc;
    const uint8_t* stage = src;                     // (partition of a multi-device encoder: the parent's staging, in place)
    if (!e->external_stage) {
      uint8_t* own = e->h_ring[n % p64b_enc::NSTAGE];
      if (src != own) copy_frames(e, own, src, (size_t)e->S * e->src_bytes);
      stage = own;
    }
    step.gquant = e->st[0].gquant;                  // (only the first frame's value is used under rate control)
    if ((rc = p64b_ctx_submit_bits(e->ctx, &step, tr, stage, &e->tickets[n % p64b_enc::DEPTH]))) return rc;
    if (!e->external_stage) e->h_src = e->h_ring[(n + 1) % p64b_enc::NSTAGE];     // its last user, frame n-3, has just been collected
    e->frames_done++;
    e->current_frame += e->p.frame_skip;
    return 0;
  }
  const uint8_t* in = src;                          // the calls below are synchronous: a partition reads the parent's staging in place
  if (!e->external_stage) {
    if (src != e->h_src) copy_frames(e, e->h_src, src, (size_t)e->S * e->src_bytes);
    in = e->h_src;
  }
  if (!e->p.rate) {
    // fixed quantiser: one device step for the whole frame of every stream, then the VLC per stream
    step.gquant = e->st[0].gquant;
    if ((rc = p64b_ctx_encode_frames(e->ctx, &step, in, e->h_mbs, e->h_levels))) return rc;
    parallel_streams(e, [&](int s) {
      StreamState& ss = e->st[s];
      p64b_bits_picture_header(ss.bits, tr);
      const p64b_mb* mb = e->h_mbs + (size_t)s * e->nmb;
      const int8_t* lv = e->h_levels + (size_t)s * e->nmb * P64B_LEVELS_PER_MB;
      for (int g = 0; g < e->ngob; g++) {
        p64b_bits_gob_header(ss.bits, g, ss.gquant);
        for (int m = 0; m < 33; m++, mb++, lv += P64B_LEVELS_PER_MB) p64b_bits_mb(ss.bits, m, mb, lv);
      }
    });
  } else {
    // rate control: GQUANT of GOB g depends on the bits written for GOBs < g, so quantise..reconstruct runs
    // per GOB (batched over streams); the overflow override is decided per MB during the VLC and patched
    // into the reconstruction at frame end.
    step.gquant = e->st[0].gquant;
    if ((rc = p64b_ctx_frame_begin(e->ctx, &step, in))) return rc;
    std::fill(e->overflow.begin(), e->overflow.end(), 0);
    parallel_streams(e, [&](int s) { p64b_bits_picture_header(e->st[s].bits, tr); });
    for (int g = 0; g < e->ngob; g++) {
      for (int s = 0; s < e->S; s++) {
        StreamState& ss = e->st[s];
        if (!first) ss.gquant = execute_quantization(e, ss, g, 0);     // p64.c:697-702
        e->quant[s] = (uint8_t)ss.gquant;
      }
      if ((rc = p64b_ctx_encode_gob(e->ctx, &step, g, e->quant.data(), e->h_mbs, e->h_levels))) return rc;
      parallel_streams(e, [&](int s) {
        StreamState& ss = e->st[s];
        p64b_bits_gob_header(ss.bits, g, ss.gquant);
        const p64b_mb* mb = e->h_mbs + (size_t)s * 33;
        const int8_t* lv = e->h_levels + (size_t)s * 33 * P64B_LEVELS_PER_MB;
        for (int m = 0; m < 33; m++, mb++, lv += P64B_LEVELS_PER_MB) {
          if (buffer_contents(e, ss, g, m) > buffer_size(e)) {          // p64.c:776-783
            p64b_mb o{}; o.mtype = 4; o.cbp = 0x3f; o.quant = (uint8_t)ss.gquant;
            p64b_bits_mb(ss.bits, m, &o, lv);
            e->overflow[(size_t)s * e->nmb + g * 33 + m] = 1;
            ss.overflows++;
          } else {
            p64b_bits_mb(ss.bits, m, mb, lv);
          }
        }
      });
    }
    if ((rc = p64b_ctx_frame_end(e->ctx, e->overflow.data()))) return rc;
  }
  for (auto& ss : e->st) {                          // p64.c:654-681
    const int64_t before = ss.total_bits;
    ss.total_bits = p64b_bits_tell(ss.bits);
    ss.last_bits = ss.total_bits - before;
    if (first) ss.first_frame_bits = ss.total_bits;
    if (e->p.rate) ss.buffer_contents_at_end = buffer_contents(e, ss, e->ngob, 0);   // p64.c:663, before 670-680
    if (e->p.rate) {
      if (first) ss.buffer_offset = buffer_size(e) / 2 - buffer_contents(e, ss, e->ngob, 0);
      ss.buffer_offset -= (int32_t)((int64_t)e->p.rate * e->p.frame_skip * e->p.frame_rate_div) / e->p.frame_rate;   // the product wraps in C int BEFORE the division (p64.c:677)
    }
  }
  e->frames_done++;
  e->current_frame += e->p.frame_skip;
  return 0;
}

int p64b_enc_finish(p64b_enc* e) {
  if (!e) return P64B_EINVAL;
  if (e->finished) return 0;
  if (!e->workers.empty()) {
    const int rc = run_all(e, Worker::FINISH, nullptr, 0);
    if (!rc) e->finished = true;
    return rc;
  }
  if (e->device_vlc)
    for (int k = e->harvested; k < e->frames_done; k++) { int rc = harvest(e, k); if (rc) return rc; }
  // p64.c:600-605: limit file growth, trailing picture header, pad with 1-bits
  int last_plus_1 = e->p.last_frame > 0 ? e->p.last_frame
                                        : e->p.start_frame + (e->frames_done ? (e->frames_done - 1) * e->p.frame_skip : 0) + 1;
  int cf = e->frames_done ? std::min(e->current_frame, last_plus_1) : e->current_frame;
  for (auto& ss : e->st) {
    if (e->device_vlc && ss.carry_len) p64b_bits_put(ss.bits, ss.carry >> (32 - ss.carry_len), (int)ss.carry_len);   // the device's pending bits
    p64b_bits_picture_header(ss.bits, cf % 32);
    p64b_bits_finish(ss.bits);
    if (e->device_vlc) {
      size_t n = 0;
      const uint8_t* t = p64b_bits_data(ss.bits, &n);
      ss.dev_bytes.insert(ss.dev_bytes.end(), t, t + n);
    }
  }
  e->finished = true;
  return 0;
}

const uint8_t* p64b_enc_data(const p64b_enc* e, int stream, size_t* nbytes) {
  if (!e || stream < 0 || stream >= e->S) { if (nbytes) *nbytes = 0; return nullptr; }
  e = route(e, &stream);
  if (e->device_vlc) { if (nbytes) *nbytes = e->st[stream].dev_bytes.size(); return e->st[stream].dev_bytes.data(); }
  return p64b_bits_data(e->st[stream].bits, nbytes);
}
p64b_ctx* p64b_enc_ctx(p64b_enc* e) { return !e ? nullptr : (e->workers.empty() ? e->ctx : e->workers.front()->kid->ctx); }
uint8_t* p64b_enc_staging(p64b_enc* e) { return e ? e->h_src : nullptr; }
int p64b_enc_partitions(const p64b_enc* e) { return !e ? P64B_EINVAL : (e->workers.empty() ? 1 : (int)e->workers.size()); }
int p64b_enc_partition(const p64b_enc* e, int k, int* device, int* first_stream, int* n_streams) {
  if (!e || k < 0 || k >= p64b_enc_partitions(e)) { p64b::set_error("bad arguments"); return P64B_EINVAL; }
  const p64b_enc* kid = e->workers.empty() ? e : e->workers[k]->kid;
  if (device) *device = kid->p.device;
  if (first_stream) *first_stream = e->workers.empty() ? 0 : e->workers[k]->first_stream;
  if (n_streams) *n_streams = kid->S;
  return 0;
}
int64_t p64b_enc_overflows(const p64b_enc* e, int stream) {
  if (!e || stream < 0 || stream >= e->S) return -1;
  e = route(e, &stream);
  return e->st[stream].overflows;
}
int p64b_enc_frame_counters(const p64b_enc* e, int stream, p64b_frame_counters* out) {
  if (!e || !out || stream < 0 || stream >= e->S) { p64b::set_error("bad arguments"); return P64B_EINVAL; }
  e = route(e, &stream);
  if (e->device_vlc) { p64b::set_error("frame counters need host_vlc = 1"); return P64B_EINVAL; }
  const StreamState& ss = e->st[stream];
  p64b_bits_counters(ss.bits, out);
  out->total_bits = (int32_t)ss.total_bits; out->last_bits = (int32_t)ss.last_bits;
  out->buffer_contents = (int32_t)ss.buffer_contents_at_end; out->buffer_size = e->p.rate / 4;
  return 0;
}
int64_t p64b_enc_first_frame_bits(const p64b_enc* e, int stream) {
  if (!e || stream < 0 || stream >= e->S) return -1;
  e = route(e, &stream);
  return e->st[stream].first_frame_bits;
}

}  // extern "C"
