// Host side of the drop-in: H.261 picture / GOB / macroblock headers and the run-level VLC.
// This is the sequential part the north star keeps on the CPU; it consumes the per-macroblock records
// and zig-zag levels the device returns and must reproduce the reference's bit stream bit for bit.
//   WritePictureHeader   marker.c:103-137        WriteGOBHeader  marker.c:182-209
//   WriteMBHeader        marker.c:288-354        EncodeDC/EncodeAC/CBPEncodeAC  codec.c:96-205, 346-355
//   Encode               huffman.c:256-284       mputv/mwtell/mwclose  stream.c:142-238
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/p64_b200.h"
#include "vlc_dev.h"
#include "vlc_tables.h"

namespace p64b {

struct Code { uint16_t bits; uint8_t len; };

static Code parse(const char* s) {
  Code c{0, 0};
  for (; *s; ++s) { c.bits = (uint16_t)((c.bits << 1) | (*s == '1')); c.len++; }
  return c;
}

struct Tables {
  Code mba[36], mtype[10], mvd[32], cbp[64];
  Code tcoef[32][16];   // [run][|level|], len 0 = escape
  Code eob, esc, first01;
  Tables() {
    memset(this, 0, sizeof(*this));
    for (auto& e : kMbaCodes) mba[e.value] = parse(e.bits);
    for (auto& e : kMtypeCodes) mtype[e.value] = parse(e.bits);
    for (auto& e : kMvdCodes) mvd[e.value] = parse(e.bits);
    for (auto& e : kCbpCodes) cbp[e.value] = parse(e.bits);
    for (auto& e : kTcoefCodes) tcoef[e.run][e.level] = parse(e.bits);
    eob = parse(kTcoefEob); esc = parse(kTcoefEscape); first01 = parse(kTcoefFirst01);
  }
};
static const Tables& T() { static const Tables t; return t; }

// MType property tables (p64.c:217-222)
static const uint8_t kQuantM[10] = {0,1,0,1,0,0,1,0,0,1};
static const uint8_t kCbpM[10]   = {0,0,1,1,0,1,1,0,1,1};
static const uint8_t kIntraM[10] = {1,1,0,0,0,0,0,0,0,0};
static const uint8_t kMfM[10]    = {0,0,0,0,1,1,1,1,1,1};
static const uint8_t kTcoefM[10] = {1,1,1,1,0,1,1,0,1,1};

void fill_dev_vlc_tables(DevVlcTables* d) {
  const Tables& t = T();
  memset(d, 0, sizeof(*d));
  auto e = [](const Code& c) { return c.len ? ((uint32_t)c.len << 16) | c.bits : 0u; };
  for (int r = 0; r < 32; r++)
    for (int l = 0; l < 16; l++) d->tcoef[r * 16 + l] = e(t.tcoef[r][l]);
  for (int i = 0; i < 10; i++) d->mtype[i] = e(t.mtype[i]);
  for (int i = 0; i < 32; i++) d->mvd[i] = e(t.mvd[i]);
  for (int i = 0; i < 64; i++) d->cbp[i] = e(t.cbp[i]);
}

}  // namespace p64b

using namespace p64b;

struct p64b_bits {
  int image_type;
  std::vector<uint8_t> buf;
  uint64_t acc = 0;   // pending bits, right-aligned
  int nacc = 0;       // number of pending bits (< 8 after every put)
  // MB-header predictors (marker.c:63-68, p64.c:716)
  int last_mba = -1, last_mtype = 0, last_mvx = 0, last_mvy = 0;
  p64b_frame_counters cnt{};   // the reference's per-frame statistics counters (p64.c:198-211, 640-649)

  inline void put(uint32_t v, int n) {         // mputv, stream.c:193-205 (n <= 32)
    acc = (acc << n) | (uint64_t)(v & (n >= 32 ? 0xffffffffu : ((1u << n) - 1)));
    nacc += n;
    while (nacc >= 8) { nacc -= 8; buf.push_back((uint8_t)(acc >> nacc)); }
    acc &= (1ull << nacc) - 1;
  }
  inline void put(const Code& c) { put(c.bits, c.len); }
  int64_t tell() const { return (int64_t)buf.size() * 8 + nacc; }
};

extern "C" {

p64b_bits* p64b_bits_create(int image_type) {
  if (image_type < 0 || image_type > 2) return nullptr;
  p64b_bits* b = new p64b_bits();
  b->image_type = image_type;
  b->buf.reserve(1 << 16);
  return b;
}
void p64b_bits_destroy(p64b_bits* b) { delete b; }

void p64b_bits_picture_header(p64b_bits* b, int tr) {
  b->put(0x10, 20);                                   // PSC (marker.h:26-27)
  b->put((uint32_t)tr, 5);
  b->put(b->image_type == P64B_IT_QCIF ? 0x00 : 0x04, 6);   // PTYPE (p64.c:408-423)
  if (b->image_type == P64B_IT_NTSC) { b->put(1, 1); b->put(0x8c, 8); }   // PSPARE (p64.c:410-412)
  b->put(0, 1);
}

void p64b_bits_gob_header(p64b_bits* b, int gob, int gquant) {
  int gread = b->image_type == P64B_IT_QCIF ? (gob << 1) : gob;          // p64.c:703-712
  b->put(1, 16);                                      // GBSC (marker.h:29-30)
  b->put((uint32_t)(gread + 1), 4);
  b->put((uint32_t)gquant, 5);
  b->put(0, 1);                                       // no GSPARE
  b->last_mba = -1; b->last_mtype = 0;                // p64.c:716
}

static inline void put_tcoef(p64b_bits* b, const Tables& t, int run, int level, bool first) {
  int a = level < 0 ? -level : level;
  if (first && run == 0 && a == 1) { b->put(t.first01); b->put(level < 0, 1); return; }
  if (run < 32 && a < 16 && t.tcoef[run][a].len) { b->put(t.tcoef[run][a]); b->put(level < 0, 1); return; }
  b->put(t.esc); b->put((uint32_t)run, 6); b->put((uint32_t)(level & 0xff), 8);   // codec.c:113-115
}

void p64b_bits_mb(p64b_bits* b, int mdu, const p64b_mb* rec, const int8_t* levels) {
  const Tables& t = T();
  const int mt = rec->mtype;
  const int mba = mdu - b->last_mba;                  // p64.c:928
  p64b_frame_counters& cn = b->cnt;
  const int64_t start = b->tell();
  int64_t mv0 = start, mv1 = start;
  b->put(t.mba[mba]);
  b->put(t.mtype[mt]);
  if (kQuantM[mt]) b->put(rec->quant, 5);
  if (kMfM[mt]) {                                     // marker.c:310-338
    mv0 = b->tell();
    int h = rec->mvx, v = rec->mvy;
    if (!kMfM[b->last_mtype] || mba != 1 || b->last_mba == -1 || b->last_mba == 10 || b->last_mba == 21) {
      b->put(t.mvd[h & 31]); b->put(t.mvd[v & 31]);
    } else {
      int dh = h - b->last_mvx, dv = v - b->last_mvy;
      if (dh < -16) dh += 32;
      if (dh > 15) dh -= 32;
      if (dv < -16) dv += 32;
      if (dv > 15) dv -= 32;
      b->put(t.mvd[dh & 31]); b->put(t.mvd[dv & 31]);
    }
    b->last_mvx = h; b->last_mvy = v;
    mv1 = b->tell();
  } else {
    b->last_mvx = b->last_mvy = 0;
  }
  if (kCbpM[mt]) b->put(t.cbp[rec->cbp]);
  b->last_mba = mdu;
  b->last_mtype = mt;
  cn.mv_bits += (int32_t)(mv1 - mv0);                 // MotionVectorBits, marker.c:344
  cn.mb_attribute_bits += (int32_t)(b->tell() - start);   // MacroAttributeBits, marker.c:354
  cn.q_use++; cn.q_sum += rec->quant;                 // p64.c:804-805
  if (mt < 10) cn.macro_type_freq[mt]++;              // p64.c:806-807
  if (!kTcoefM[mt]) return;
  for (int c = 0; c < 6; c++) {                       // p64.c:931-950
    if (!(rec->cbp & (1 << (5 - c)))) continue;
    const int8_t* l = levels + 64 * c;
    int k = 0;
    bool first = false;
    (c < 4 ? cn.y_type_freq : cn.uv_type_freq)[mt]++;     // p64.c:937-938
    const int64_t bstart = b->tell();
    if (kCbpM[mt]) first = true;                      // CBPEncodeAC(0,.)  codec.c:140-205
    else {                                            // EncodeDC + EncodeAC(1,.)  codec.c:346-355, 96-130
      int dc = (uint8_t)l[0];
      if (dc > 254) dc = 254;
      if (dc < 1) dc = 1;
      if (dc != 1) cn.number_nz++;                     // codec.c:351
      if (dc == 128) dc = 255;
      b->put((uint32_t)dc, 8);
      k = 1;
    }
    int run = 0;
    bool any = !first;
    for (; k < 64; k++) {
      int v = l[k];
      if (!v) { run++; continue; }
      put_tcoef(b, t, run, v, first);
      first = false; any = true; run = 0;
      cn.number_nz++;                                 // codec.c:125, 169, 200
    }
    (c < 4 ? cn.y_bits : (c == 4 ? cn.u_bits : cn.v_bits)) += (int32_t)(b->tell() - bstart);   // CodedBlockBits, p64.c:949-951
    if (any) cn.eob_bits += t.eob.len;                // codec.c:129, 204
    if (any) b->put(t.eob);                           // an all-zero CBP block gets no EOB (codec.c:169-174)
  }
}

void p64b_bits_put(p64b_bits* b, uint32_t value, int nbits) { if (nbits > 0 && nbits <= 32) b->put(value, nbits); }
int64_t p64b_bits_tell(const p64b_bits* b) { return b->tell(); }

size_t p64b_bits_finish(p64b_bits* b) {
  while (b->nacc) b->put(1, 1);                       // mwclose pads with ones
  return b->buf.size();
}
const uint8_t* p64b_bits_data(const p64b_bits* b, size_t* n) {
  if (n) *n = b->buf.size();
  return b->buf.data();
}
void p64b_bits_counters(const p64b_bits* b, p64b_frame_counters* out) { if (b && out) *out = b->cnt; }
void p64b_bits_counters_reset(p64b_bits* b) { if (b) b->cnt = p64b_frame_counters{}; }   // p64.c:640-649
void p64b_bits_reset(p64b_bits* b) {
  b->cnt = p64b_frame_counters{};
  b->buf.clear(); b->acc = 0; b->nacc = 0; b->last_mba = -1; b->last_mtype = 0; b->last_mvx = b->last_mvy = 0;
}

}  // extern "C"
