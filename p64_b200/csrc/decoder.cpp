// Decoder: the sequential half -- an H.261 bit-stream parser restating the reference's
//   p64DecodeSequence / p64DecodeGOB / DecompressMDU   p64.c:1022-1237
//   ReadHeaderHeader / ReadHeaderTrailer / ReadPictureHeader / ReadGOBHeader / ReadMBHeader   marker.c:144-450
//   DecodeDC / DecodeAC / CBPDecodeAC   codec.c:214-340        Decode   huffman.c:292-340
// -- stays on the host and turns each picture into the macroblock records + levels that the device's inverse half
// (mb_decode_kernel through p64b_ctx_decode_frames) reconstructs.  The code tables are the same bit strings the encoder's
// writers use (vlc_tables.h).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/p64_b200.h"
#include "vlc_tables.h"

namespace p64b { void set_error(const std::string& s); }
using namespace p64b;

namespace {

constexpr int kEscape = 0x7fff, kEob = 0;

struct BitReader {                       // mgetb / mgetv, stream.c:170-230 (MSB first)
  const uint8_t* d = nullptr;
  size_t nbits = 0, pos = 0;
  bool eof = false;
  inline uint32_t peek(int n) const {    // n <= 24; zeros beyond the end
    uint32_t v = 0;
    const size_t byte = pos >> 3;
    for (int i = 0; i < 4; i++) v = (v << 8) | (byte + i < (nbits >> 3) ? d[byte + i] : 0u);
    return (v << (pos & 7)) >> (32 - n);
  }
  inline void skip(int n) { pos += n; if (pos > nbits) { pos = nbits; eof = true; } }
  inline uint32_t get(int n) { const uint32_t v = peek(n); skip(n); return v; }
};

struct DecTable {                        // prefix-code lookup by the next `maxlen` bits
  int maxlen = 0;
  std::vector<int32_t> lut;              // (value << 8) | length, -1 = no such code ("Invalid State Reached", huffman.c:311)
  void build(const std::vector<std::pair<const char*, int>>& codes) {
    for (auto& c : codes) maxlen = std::max(maxlen, (int)strlen(c.first));
    lut.assign((size_t)1 << maxlen, -1);
    for (auto& c : codes) {
      const int len = (int)strlen(c.first);
      uint32_t bits = 0;
      for (const char* s = c.first; *s; ++s) bits = (bits << 1) | (*s == '1');
      const uint32_t lo = bits << (maxlen - len), hi = lo + (1u << (maxlen - len));
      for (uint32_t i = lo; i < hi; i++) lut[i] = (c.second << 8) | len;
    }
  }
  inline int decode(BitReader& b, bool* bad) const {
    const int32_t e = lut[b.peek(maxlen)];
    if (e < 0) { *bad = true; return 0; }
    b.skip(e & 0xff);
    return e >> 8;
  }
};

struct DecTables {
  DecTable mba, mtype, mvd, cbp, t1, t2;   // MBADHuff, T3DHuff, MVDDHuff, CBPDHuff, T1DHuff, T2DHuff (huffman.c:102-135)
  DecTables() {
    std::vector<std::pair<const char*, int>> v;
    for (auto& e : kMbaCodes) v.push_back({e.bits, e.value});
    mba.build(v); v.clear();
    for (auto& e : kMtypeCodes) v.push_back({e.bits, e.value});
    mtype.build(v); v.clear();
    for (auto& e : kMvdCodes) v.push_back({e.bits, e.value});
    mvd.build(v); v.clear();
    for (auto& e : kCbpCodes) v.push_back({e.bits, e.value});
    cbp.build(v); v.clear();
    // T1: EOB "10", (0,1) "11"; T2 (first coefficient of a CBP-coded block): (0,1) "1", no EOB; value = level | run << 8
    std::vector<std::pair<const char*, int>> a, b;
    a.push_back({kTcoefEob, kEob});
    a.push_back({kTcoefEscape, kEscape}); b.push_back({kTcoefEscape, kEscape});
    for (auto& e : kTcoefCodes) {
      const int val = e.level | (e.run << 8);
      a.push_back({e.bits, val});
      b.push_back({(e.run == 0 && e.level == 1) ? kTcoefFirst01 : e.bits, val});
    }
    t1.build(a); t2.build(b);
  }
};
const DecTables& DT() { static const DecTables t; return t; }

const uint8_t kQuantM[10] = {0,1,0,1,0,0,1,0,0,1}, kCbpM[10] = {0,0,1,1,0,1,1,0,1,1};       // p64.c:217-222
const uint8_t kMfM[10] = {0,0,0,0,1,1,1,1,1,1}, kTcoefM[10] = {1,1,1,1,0,1,1,0,1,1}, kIntraM[10] = {1,1,0,0,0,0,0,0,0,0};

inline int sext5(int v) { return (v & 0x10) ? (v | ~0x1f) : v; }     // bit_set_mask[4] / extend_mask[4], marker.c:410-413

}  // namespace

struct p64b_parser {
  BitReader br;
  int image_type = -1, ngob = 0, nmb = 0;
  bool ended = false;                 // nothing more to decode
  bool at_gob = false;                // a GBSC has been consumed and its 4-bit trailer is next (ReadHeaderTrailer)
  int tr = 0;                         // TemporalReference of the picture being decoded
  int current_frame = 0, temporal_offset = 0;
  // decoder globals that persist across macroblocks / GOBs (p64.c:85-98)
  int gquant = 0, mquant = 0, mtype = 0, mvdh = 0, mvdv = 0;
  std::string error;

  bool read_header_header() {         // ReadHeaderHeader, marker.c:246-262: 16 bits GBSC
    if (br.pos + 16 > br.nbits) { br.eof = true; return false; }
    return br.get(16) == 1;
  }
  void read_picture_header() {        // ReadPictureHeader, marker.c:144-172 (after PSC): TR, PTYPE, PEI/PSPARE
    tr = (int)br.get(5);
    ptype = (int)br.get(6);
    pspare_enable = false;
    while (br.get(1) && !br.eof) { pspare_enable = true; pspare = (int)br.get(8); }
  }
  int ptype = 0, pspare = 0;
  bool pspare_enable = false;

  // one block's coefficients in transmission order (codec.c:214-340); false on a corrupt stream
  bool decode_block(bool cbp_type, int8_t* lv) {
    const DecTables& t = DT();
    bool bad = false;
    int k = 0;
    auto one = [&](const DecTable& tab, bool* eob) -> bool {
      int r = tab.decode(br, &bad), l;
      if (bad) return false;
      if (r == kEob) { *eob = true; return true; }
      if (r == kEscape) { r = (int)br.get(6); l = (int)br.get(8); }
      else { l = r & 0xff; r >>= 8; if (br.get(1)) l = -l; }
      l = (int)(int8_t)(l & 0xff);                      // bit_set_mask[7] / extend_mask[7]
      k += r;
      if (k > 63) return false;                         // the reference would write past its matrix here
      lv[k++] = (int8_t)l;
      return true;
    };
    bool eob = false;
    if (cbp_type) {                                     // CBPDecodeAC(0, .): the first code cannot be EOB
      if (!one(t.t2, &eob)) return false;
    } else {                                            // DecodeDC + DecodeAC(1, .)
      int l = (int)br.get(8);
      if (l == 255) l = 128;
      lv[0] = (int8_t)(uint8_t)l;
      k = 1;
    }
    while (k < 64) {
      if (!one(t.t1, &eob)) return false;
      if (eob) return true;
    }
    (void)t.t1.decode(br, &bad);                        // "EOB expected" (codec.c:245-249): consumed whatever it is
    return !bad;
  }
};

extern "C" {

int p64b_parser_create(p64b_parser** out, const uint8_t* data, size_t nbytes) {
  if (!out || !data) { set_error("NULL argument"); return P64B_EINVAL; }
  p64b_parser* p = new p64b_parser();
  p->br.d = data; p->br.nbits = nbytes * 8;
  // p64DecodeSequence, p64.c:1030-1035 + the first pass through the loop (1038-1044, 1067-1112)
  if (!p->read_header_header()) { delete p; set_error("Illegal GOB Start Code at the start of the stream"); return P64B_EIO; }
  if ((int)p->br.get(4) - 1 >= 0) { delete p; set_error("stream does not start with a picture header"); return P64B_EIO; }
  p->read_picture_header();
  if (p->ptype & 0x04) p->image_type = (p->pspare_enable && p->pspare == 0x8c) ? P64B_IT_NTSC : P64B_IT_CIF;   // p64.c:1071-1081
  else p->image_type = P64B_IT_QCIF;
  p->ngob = p64b_num_gob(p->image_type); p->nmb = p64b_num_mb(p->image_type);
  p->temporal_offset = (p->tr - p->current_frame) % 32;
  p->at_gob = p->read_header_header();
  p->ended = !p->at_gob;
  *out = p;
  return 0;
}

void p64b_parser_destroy(p64b_parser* p) { delete p; }
int p64b_parser_image_type(const p64b_parser* p) { return p ? p->image_type : P64B_EINVAL; }

int p64b_parser_next_picture(p64b_parser* p, p64b_mb* mbs, int8_t* levels, int* temporal_reference, int* repeat) {
  if (!p || !mbs || !levels) { set_error("NULL argument"); return P64B_EINVAL; }
  if (p->ended) return 0;
  const DecTables& t = DT();
  BitReader& br = p->br;
  memset(mbs, 0, (size_t)p->nmb * sizeof(p64b_mb));
  memset(levels, 0, (size_t)p->nmb * P64B_LEVELS_PER_MB);
  if (temporal_reference) *temporal_reference = p->tr;
  bool end_frame = false;               // EndFrame: the stream ended inside the picture
  for (;;) {
    int gread = -1;
    if (!end_frame) gread = (int)br.get(4) - 1;             // ReadHeaderTrailer
    if (gread < 0 || end_frame) {                           // end of this picture
      if (!end_frame) p->read_picture_header(); else p->tr = (p->tr + 1) & 31;
      int cnt = 0;                                          // p64.c:1047-1054: temporal-reference gaps repeat the picture
      while (((p->current_frame + p->temporal_offset) % 32) != p->tr && cnt < 64) { cnt++; p->current_frame++; }
      if (repeat) *repeat = cnt;
      p->at_gob = !end_frame && p->read_header_header();
      p->ended = !p->at_gob;
      return 1;
    }
    // ---- p64DecodeGOB, p64.c:1134-1170
    p->gquant = (int)br.get(5);                             // ReadGOBHeader, marker.c:271-281
    while (br.get(1) && !br.eof) (void)br.get(8);           // GSPARE
    const int gob = p->image_type == P64B_IT_QCIF ? (gread >> 1) : gread;
    if (gob >= p->ngob) { end_frame = true; continue; }     // "Buffer Overflow: Current:%d Number:%d"
    int last_mba = -1;
    for (;;) {                                              // ReadMBHeader, marker.c:374-450
      bool bad = false;
      int mba;
      do { mba = t.mba.decode(br, &bad); } while (mba == 34 && !bad && !br.eof);     // stuffing
      if (bad || br.eof) { end_frame = true; break; }
      if (mba == 35) break;                                 // start code: the next header follows
      const int last_mtype = p->mtype;
      p->mtype = t.mtype.decode(br, &bad);
      if (bad) { end_frame = true; break; }
      const int mt = p->mtype;
      if (kQuantM[mt]) p->mquant = (int)br.get(5);
      if (kMfM[mt]) {
        const int rh = sext5(t.mvd.decode(br, &bad)), rv = sext5(t.mvd.decode(br, &bad));
        if (!kMfM[last_mtype] || mba != 1 || last_mba == -1 || last_mba == 10 || last_mba == 21) { p->mvdh = rh; p->mvdv = rv; }
        else {
          p->mvdh += rh; p->mvdv += rv;
          if (p->mvdh < -16) p->mvdh += 32;
          if (p->mvdh > 15) p->mvdh -= 32;
          if (p->mvdv < -16) p->mvdv += 32;
          if (p->mvdv > 15) p->mvdv -= 32;
        }
      } else {
        p->mvdh = p->mvdv = 0;
      }
      int cbp = 0x3f;
      if (kCbpM[mt]) cbp = t.cbp.decode(br, &bad);
      if (bad) { end_frame = true; break; }
      // ---- DecompressMDU, p64.c:1179-1237
      last_mba += mba;
      if (last_mba >= 33) { end_frame = true; break; }      // "Apparent MDU out of range" / end of file
      if (kMfM[mt]) {
        // A vector that takes the 16x16 prediction outside the picture cannot come from a conforming encoder (the reference
        // decoder would read outside its planes, io.c:200-313): the stream is corrupt from here on.
        const int m = last_mba, W = p64b_width(p->image_type), H = p64b_height(p->image_type);
        const int col = p->image_type == P64B_IT_QCIF ? m % 11 : (gob & 1) * 11 + m % 11;
        const int row = p->image_type == P64B_IT_QCIF ? gob * 3 + m / 11 : (gob >> 1) * 3 + m / 11;
        const int px = col * 16 + p->mvdh, py = row * 16 + p->mvdv;
        if (px < 0 || py < 0 || px > W - 16 || py > H - 16) { end_frame = true; break; }
      }
      int use_quant = p->gquant;
      if (kQuantM[mt]) { use_quant = p->mquant; p->gquant = p->mquant; }
      p64b_mb& r = mbs[gob * 33 + last_mba];
      r.mtype = (uint8_t)mt; r.cbp = (uint8_t)cbp; r.mvx = (int8_t)p->mvdh; r.mvy = (int8_t)p->mvdv;
      r.quant = (uint8_t)use_quant; r.reserved = 1;
      int8_t* lv = levels + (size_t)(gob * 33 + last_mba) * P64B_LEVELS_PER_MB;
      memset(lv, 0, P64B_LEVELS_PER_MB);                    // a macroblock sent twice: the last one counts
      bool ok = true;
      if (kTcoefM[mt])
        for (int c = 0; c < 6 && ok; c++)
          if (cbp & (1 << (5 - c))) ok = p->decode_block(kCbpM[mt] != 0, lv + 64 * c);
      if (!ok || br.eof) { end_frame = true; break; }
    }
  }
}

}  // extern "C"

// ---- the decoder object: parser + device context ----------------------------------------------------------------

struct p64b_dec {
  p64b_parser* parser = nullptr;
  p64b_ctx* ctx = nullptr;
  std::vector<p64b_mb> mbs;
  std::vector<int8_t> levels;
};

extern "C" {

int p64b_dec_create(p64b_dec** out, int device, const uint8_t* data, size_t nbytes) {
  if (!out) { set_error("NULL argument"); return P64B_EINVAL; }
  p64b_dec* d = new p64b_dec();
  int rc = p64b_parser_create(&d->parser, data, nbytes);
  if (!rc) rc = p64b_ctx_create(&d->ctx, device, d->parser->image_type, 1);     // MakeIob / InitFS / ClearFS, p64.c:1108-1110
  if (rc) { p64b_dec_destroy(d); return rc; }
  d->mbs.resize(d->parser->nmb);
  d->levels.resize((size_t)d->parser->nmb * P64B_LEVELS_PER_MB);
  *out = d;
  return 0;
}

void p64b_dec_destroy(p64b_dec* d) {
  if (!d) return;
  p64b_parser_destroy(d->parser);
  p64b_ctx_destroy(d->ctx);
  delete d;
}

int p64b_dec_image_type(const p64b_dec* d) { return d ? d->parser->image_type : P64B_EINVAL; }

int p64b_dec_next_picture(p64b_dec* d, uint8_t* yuv, int* repeat) {
  if (!d || !yuv) { set_error("NULL argument"); return P64B_EINVAL; }
  int tr = 0;
  int rc = p64b_parser_next_picture(d->parser, d->mbs.data(), d->levels.data(), &tr, repeat);
  if (rc != 1) return rc;
  if ((rc = p64b_ctx_decode_frames(d->ctx, d->mbs.data(), d->levels.data()))) return rc;
  if ((rc = p64b_ctx_download_recon(d->ctx, 0, yuv))) return rc;
  return 1;
}

}  // extern "C"
