// p64b -- command line with the reference's flags (p64.c:262-449, Help() p64.c:1537-1568) on top of the B200 hot path.
// Encoder, and with -d the decoder (host parser + the device's inverse half); video I/O is Y4M (the reference's
// raw-component path crashes at end of sequence).  Extra flags that the reference lacks are spelled with two dashes:
//   --me tss|full   the stock three-step search (me.c:352) or the exhaustive FastBME (me.c:351)   [default tss]
//   --intra-only    stand-in for `-o < test.intra` (every macroblock intra)
//   --device N      CUDA device
// Several file prefixes on one command line are encoded together as one batch of independent streams (Prefix.y4m ->
// Prefix.p64 each; same picture type and flags for all) -- the batch is what fills a GPU, one CIF stream cannot.
// Same stdout markers as the reference (START>SEQUENCE, START>Frame: n, END>Frame: n, END>SEQUENCE, and the
// "Bits for first frame" line, p64.c:595-611, 627, 665).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/p64_b200.h"

static void help() {
  printf("p64b -a StartFrame -b LastFrame [-NTSC] [-CIF] [-QCIF] [-y4m]\n"
         "     [-f FrameRate[/Div]] [-i SearchLimit] [-k FrameSkip] [-q Quantization] [-r Rate] [-x FileSizeBits]\n"
         "     [-s StreamFile] [-l 1] [-d] [--me tss|full] [--intra-only] [--device N | --devices 0-7 [--balance-links]] Y4MFilePrefix [MorePrefixes...]\n"
         "Encodes PrefixYUV4MPEG2 file `Prefix.y4m` (or `-` for stdin) into an H.261 stream; the data-parallel hot path\n"
         "(motion estimation, DCT, quantisation, reconstruction) runs on the GPU. There is no CPU fallback.\n");
}

// PrintFrameStatistics (p64.c:1299-1332) + Statistics (stat.c:52-63): the block `-l 1` prints after every frame; the plane
// statistics come from the device (p64b_ctx_statistics), the bit counters from the host bit writer.
static void print_frame_statistics(p64b_enc* enc, const p64b_enc_params& p) {
  p64b_frame_counters c;
  if (p64b_enc_frame_counters(enc, 0, &c)) return;
  const int nmb = p64b_num_mb(p.image_type);
  printf("Total No of Bits: %8d  Bits for Frame: %8d\n", c.total_bits, c.last_bits);
  if (p.rate) printf("Buffer Contents: %8d  out of: %8d\n", c.buffer_contents, c.buffer_size);
  printf("MB Attribute Bits: %6d  MV Bits: %6d   EOB Bits: %6d\n", c.mb_attribute_bits, c.mv_bits, c.eob_bits);
  printf("Y Bits: %7d  U Bits: %7d  V Bits: %7d  Total Bits: %7d\n", c.y_bits, c.u_bits, c.v_bits, c.y_bits + c.u_bits + c.v_bits);
  printf("MV StepSize: %f  MV NumberNonZero: %f  MV NumberZero: %f\n", (double)c.q_sum / (double)c.q_use,
         (double)c.number_nz / (double)(nmb * 6), (double)(nmb * 6 * 64 - c.number_nz) / (double)(nmb * 6));
  printf("Code MType: ");
  for (int x = 0; x < 10; x++) printf("%5d", x);
  printf("\nMacro Freq: ");
  for (int x = 0; x < 10; x++) printf("%5d", c.macro_type_freq[x]);
  printf("\nY     Freq: ");
  for (int x = 0; x < 10; x++) printf("%5d", c.y_type_freq[x]);
  printf("\nUV    Freq: ");
  for (int x = 0; x < 10; x++) printf("%5d", c.uv_type_freq[x]);
  printf("\n");
  p64b_plane_stats sums[3];
  if (p64b_ctx_statistics(p64b_enc_ctx(enc), sums)) return;
  for (int i = 0; i < 3; i++) {
    p64b_stat s;
    p64b_stat_from_sums(&sums[i], &s);
    printf("Comp: %d  MRSNR: %2.2f  SNR: %2.2f  PSNR: %2.2f  MSE: %4.2f  Entropy: %1.2f\n", i, s.mrsnr, s.snr, s.psnr, s.mse, s.entropy);
  }
}

// -d: p64DecodeSequence (p64.c:1022-1126) -- StreamFile -> Prefix.y4m with the reference's Y4M header (p64.c:249-251, 1104)
// and its START>Frame / END> Frame markers.
static int decode_stream(const std::string& stream_file, const std::string& prefix, const p64b_enc_params& p) {
  FILE* f = fopen(stream_file.c_str(), "rb");
  if (!f) { printf("Cannot Open Input File\n"); return 1; }
  std::vector<uint8_t> data;
  uint8_t buf[65536];
  for (size_t n; (n = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + n);
  fclose(f);
  p64b_dec* dec = nullptr;
  if (p64b_dec_create(&dec, p.device, data.data(), data.size())) { fprintf(stderr, "p64b: %s\n", p64b_last_error()); return 2; }
  const int it = p64b_dec_image_type(dec);
  FILE* out = fopen((prefix + ".y4m").c_str(), "wb");
  if (!out) { printf("Cannot Open Output File\n"); return 1; }
  fprintf(out, "YUV4MPEG2 W%d H%d C420jpeg Ip F%i:%i\n", p64b_width(it), p64b_height(it), p.frame_rate, p.frame_rate_div);
  std::vector<uint8_t> frame(p64b_frame_bytes(it));
  int cf = 0, rep = 0, rc;
  printf("START>Frame: %d\n", cf);
  while ((rc = p64b_dec_next_picture(dec, frame.data(), &rep)) == 1) {
    for (int i = 0; i < rep; i++) {
      printf("END> Frame: %d\n", cf++);
      fwrite("FRAME\n", 1, 6, out);
      fwrite(frame.data(), 1, frame.size(), out);
    }
    printf("START>Frame: %d\n", cf);
  }
  fclose(out);
  p64b_dec_destroy(dec);
  if (rc < 0) { fprintf(stderr, "p64b: %s\n", p64b_last_error()); return 2; }
  return 0;
}

// Several prefixes: one batch of independent streams, one frame of every stream per device step (p64b_enc_* with
// n_streams = number of files).  Every file must have the selected picture size and the same chroma type.
static int encode_batch(const std::vector<std::string>& prefixes, p64b_enc_params p, int start, int last) {
  const int S = (int)prefixes.size();
  std::vector<p64b_y4m*> in(S, nullptr);
  p64b_y4m_info info0{};
  for (int s = 0; s < S; s++) {
    const std::string path = prefixes[s] + ".y4m";
    if (p64b_y4m_open(&in[s], path.c_str())) { fprintf(stderr, "Unable to open '%s': %s\n", path.c_str(), p64b_last_error()); return -1; }
    p64b_y4m_info info;
    p64b_y4m_get_info(in[s], &info);
    if (s == 0) info0 = info;
    if (info.width != p64b_width(p.image_type) || info.height != p64b_height(p.image_type) || info.chroma != info0.chroma) {
      fprintf(stderr, "p64b: '%s' does not match the selected image type / the first file's chroma type\n", path.c_str());
      return 3;
    }
  }
  p.input_chroma = info0.chroma;
  p.n_streams = S;
  p64b_enc* enc = nullptr;
  if (p64b_enc_create(&enc, &p)) { fprintf(stderr, "p64b: %s\n", p64b_last_error()); return 2; }
  const size_t fb = (size_t)info0.frame_bytes;
  for (int i = start; i > 0; --i)
    for (int s = 0; s < S; s++)
      if (p64b_y4m_read_frame(in[s], p64b_enc_staging(enc) + s * fb) != 1) return 3;
  printf("START>SEQUENCE\n");
  for (int cf = start; cf <= last; cf += p.frame_skip) {
    printf("START>Frame: %d\n", cf);
    uint8_t* frame = p64b_enc_staging(enc);
    for (int s = 0; s < S; s++)
      if (p64b_y4m_read_frame(in[s], frame + s * fb) != 1) { p64b_enc_destroy(enc); return 3; }
    if (p64b_enc_encode(enc, frame)) { fprintf(stderr, "p64b: %s\n", p64b_last_error()); return 2; }
    printf("END>Frame: %d\n", cf);
  }
  p64b_enc_finish(enc);
  for (int s = 0; s < S; s++) {
    size_t n = 0;
    const uint8_t* data = p64b_enc_data(enc, s, &n);
    FILE* out = fopen((prefixes[s] + ".p64").c_str(), "wb");
    if (!out || fwrite(data, 1, n, out) != n) { printf("Cannot Open Output File\n"); return 1; }
    fclose(out);
    printf("%s: Bits for first frame: %lld   Number of buffer overflows: %lld\n", prefixes[s].c_str(),
           (long long)p64b_enc_first_frame_bits(enc, s), (long long)p64b_enc_overflows(enc, s));
    p64b_y4m_close(in[s]);
  }
  printf("END>SEQUENCE\n");
  p64b_enc_destroy(enc);
  return 0;
}

int main(int argc, char** argv) {
  p64b_enc_params p;
  p64b_enc_default_params(&p);
  int start = 0, last = 0, file_size_bits = 0, loud = 0;
  bool decode = false;
  std::string prefix, stream_file;
  std::vector<std::string> prefixes;
  if (argc == 1) { help(); return -1; }
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { help(); exit(1); } return argv[++i]; };
    if (a == "-NTSC") p.image_type = P64B_IT_NTSC;
    else if (a == "-CIF") p.image_type = P64B_IT_CIF;
    else if (a == "-QCIF") p.image_type = P64B_IT_QCIF;
    else if (a == "-y4m") {}
    else if (a == "-d") decode = true;                     // p64.c:307-309
    else if (a == "--me") { std::string m = next(); p.me_mode = m == "full" ? P64B_ME_FULL : P64B_ME_TSS; }
    else if (a == "--intra-only" || a == "-o") p.force_intra = 1;
    else if (a == "--device") p.device = atoi(next());
    else if (a == "--balance-links") p.balance_links = 1;      // with --devices: blocks in proportion to each GPU's host-link share
    else if (a == "--devices") {                           // "0-7" or "0,2,3": streams (prefixes) are partitioned over these GPUs
      p.n_devices = 0;
      for (const char* v = next(); *v;) {
        char* end;
        const int lo = (int)strtol(v, &end, 10);
        int hi = lo;
        if (end == v) { printf("Bad device list.\n"); return 3; }
        if (*end == '-') { const char* q = end + 1; hi = (int)strtol(q, &end, 10); if (end == q) { printf("Bad device list.\n"); return 3; } }
        for (int d = lo; d <= hi && p.n_devices < P64B_MAX_DEVICES; d++) p.devices[p.n_devices++] = d;
        v = *end == ',' ? end + 1 : end;
        if (*end && *end != ',') { printf("Bad device list.\n"); return 3; }
      }
    }
    else if (a == "-a") start = atoi(next());
    else if (a == "-b") last = atoi(next());
    else if (a == "-f") {                                  // p64.c:316-340
      const char* v = next();
      p.frame_rate = atoi(v); p.frame_rate_div = 1;
      if (const char* s = strpbrk(v, "/:")) p.frame_rate_div = atoi(s + 1) > 0 ? atoi(s + 1) : 1;
      else if (const char* d = strchr(v, '.')) {
        for (size_t k = strlen(d + 1); k > 0; --k) p.frame_rate_div *= 10;
        p.frame_rate = p.frame_rate * p.frame_rate_div + atoi(d + 1);
      }
    }
    else if (a == "-i") p.search_limit = atoi(next());
    else if (a == "-k") p.frame_skip = atoi(next());
    else if (a == "-q") p.initial_quant = atoi(next());
    else if (a == "-r") p.rate = atoi(next());
    else if (a == "-x") file_size_bits = atoi(next());
    else if (a == "-s") stream_file = next();
    else if (a == "-l") loud = atoi(next());               // Loud (p64.c:348-350): > 0 prints the frame statistics
    else if (a == "-z") next();                            // accepted, ignored (component file suffixes)
    else if (a == "-v" || a == "-c" || a == "-p") {}
    else if (a == "-") prefix = "-";
    else if (a[0] == '-') { printf("Illegal Option %s\n", a.c_str()); return 3; }
    else { prefix = a; prefixes.push_back(a); }
  }
  if (prefix.empty()) { printf("A file prefix should be specified.\n"); return 3; }
  if (decode) return decode_stream(stream_file.empty() ? prefix + ".p64" : stream_file, prefix, p);
  if (start > last) { printf("Need positive number of frames.\n"); return 3; }
  if (p.search_limit < 1 || p.search_limit > 31 || p.initial_quant < 0 || p.initial_quant > 31) { printf("Parameter out of bounds.\n"); return 3; }
  if (file_size_bits) {
    // p64.c:572-573 in the reference's C int: FileSizeBits*FrameRate wraps at 32 bits BEFORE the divisions (at the default
    // 30000/1001 for any FileSizeBits > 71582).  A positive wrapped rate is reproduced; a wrapped rate <= 0 (which the reference
    // runs with a negative QDFact) is refused -- a stated departure (out of scope).
    const int32_t prod = (int32_t)((int64_t)file_size_bits * p.frame_rate);
    p.rate = prod / p.frame_rate_div / (p.frame_skip * (last - start + 1));
    if (p.rate <= 0) {
      printf("-x %d: FileSizeBits*FrameRate overflows the reference's int arithmetic (Rate: %d); use -r\n", file_size_bits, p.rate);
      return 3;
    }
  }
  p.start_frame = start;
  p.last_frame = last + 1;
  if (stream_file.empty()) stream_file = prefix + ".p64";
  if (prefixes.size() > 1) return encode_batch(prefixes, p, start, last);

  // ingest: the library's Y4M reader fills the encoder's pinned staging buffer in place; chroma types other than
  // 420jpeg are converted on the device (y4m_input.c:195-545)
  p64b_y4m* in = nullptr;
  std::string path = prefix == "-" ? "-" : prefix + ".y4m";
  if (p64b_y4m_open(&in, path.c_str())) { fprintf(stderr, "Unable to open '%s': %s\n", path.c_str(), p64b_last_error()); return -1; }
  p64b_y4m_info info;
  p64b_y4m_get_info(in, &info);
  if (info.width != p64b_width(p.image_type) || info.height != p64b_height(p.image_type)) {
    fprintf(stderr, "p64b: %dx%d input does not match the selected image type (%dx%d)\n", info.width, info.height,
            p64b_width(p.image_type), p64b_height(p.image_type));
    return 3;
  }
  p.input_chroma = info.chroma;
  if (loud > 0) p.host_vlc = 1;                            // the per-category bit counters live in the host bit writer
  p64b_enc* enc = nullptr;
  if (p64b_enc_create(&enc, &p)) { fprintf(stderr, "p64b: %s\n", p64b_last_error()); return 2; }
  for (int i = start; i > 0; --i)                           // seek to StartFrame (p64.c:562-565)
    if (p64b_y4m_read_frame(in, p64b_enc_staging(enc)) != 1) return 3;
  if (p.rate && !p.initial_quant) {                        // p64.c:574-586
    int iq = 10000000 / p.rate;
    iq = iq > 31 ? 31 : (iq < 1 ? 1 : iq);
    printf("Rate: %d   QDFact: %d  QOffs: %d\n", p.rate, p.rate / 320, 1);
    printf("Starting Quantization: %d\n", iq);
  }
  printf("START>SEQUENCE\n");
  for (int cf = start; cf <= last; cf += p.frame_skip) {
    printf("START>Frame: %d\n", cf);
    uint8_t* frame = p64b_enc_staging(enc);                  // the pinned upload buffer of THIS frame (it rotates: frames are pipelined)
    if (p64b_y4m_read_frame(in, frame) != 1) { p64b_enc_destroy(enc); return 3; }
    if (p64b_enc_encode(enc, frame)) { fprintf(stderr, "p64b: %s\n", p64b_last_error()); return 2; }
    if (loud > 0) print_frame_statistics(enc, p);
    printf("END>Frame: %d\n", cf);
  }
  p64b_enc_finish(enc);
  size_t n = 0;
  const uint8_t* data = p64b_enc_data(enc, 0, &n);
  FILE* out = fopen(stream_file.c_str(), "wb");
  if (!out || fwrite(data, 1, n, out) != n) { printf("Cannot Open Output File\n"); return 1; }
  fclose(out);
  printf("END>SEQUENCE\n");
  printf("Bits for first frame: %lld   Number of buffer overflows: %lld\n", (long long)p64b_enc_first_frame_bits(enc, 0),
         (long long)p64b_enc_overflows(enc, 0));
  p64b_enc_destroy(enc);
  p64b_y4m_close(in);
  return 0;
}
