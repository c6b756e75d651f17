/* fake_scheduler -- drives the four gr-doa blocks (GPU implementations in ../lib) the way GNU Radio's thread-per-block
 * scheduler would, without GNU Radio: scheduler-owned host ring buffers, history()/forecast()/consume_each() honoured,
 * work() called with a scheduler-chosen, varying noutput_items.  Flowgraph:
 *
 *     M x vector_source_c -> autocorrelate -> MUSIC_lin_array -> find_local_max -> 2 x vector_sink_f
 *                                          \-> rootMUSIC_linear_array -> vector_sink_f
 *
 * usage: fake_scheduler <in.c64> <M> <N> <overlap> <avg> <d> <T> <P> <K> <out_prefix>
 *   in.c64: M channel streams of equal length, channel-major, raw complex64.
 *   writes <out_prefix>.R.c64, .spec.f32, .val.f32, .loc.f32, .aoa.f32 (raw) for the test suite to check.
 *   DOA_HARNESS_SC16_IN=<file of the same streams as int16 I/Q, channel-major>: a doa.music_chain made with make_sc16
 *   (scale 1/32768) sees the same scheduler calls on those items; its outputs go to <out_prefix>.sval.f32 / .sloc.f32.
 */
#include <doa/MUSIC_lin_array.h>
#include <doa/autocorrelate.h>
#include <doa/music_chain.h>
#include <doa/rootmusic_chain.h>
#include <doa/find_local_max.h>
#include <doa/rootMUSIC_linear_array.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static std::vector<char> slurp(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  std::vector<char> b(n);
  if (std::fread(b.data(), 1, n, f) != (size_t)n) std::exit(2);
  std::fclose(f);
  return b;
}
static void dump(const std::string& path, const void* p, size_t bytes) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f || std::fwrite(p, 1, bytes, f) != bytes) { std::perror(path.c_str()); std::exit(2); }
  std::fclose(f);
}

int main(int argc, char** argv) {
  if (argc != 11) { std::fprintf(stderr, "usage: %s in.c64 M N overlap avg d T P K out_prefix\n", argv[0]); return 2; }
  const int M = std::atoi(argv[2]), N = std::atoi(argv[3]), overlap = std::atoi(argv[4]), avg = std::atoi(argv[5]);
  const float d = (float)std::atof(argv[6]);
  const int T = std::atoi(argv[7]), P = std::atoi(argv[8]), K = std::atoi(argv[9]);
  const std::string prefix = argv[10];
  std::vector<char> raw = slurp(argv[1]);
  const size_t L = raw.size() / sizeof(gr_complex) / M;
  const gr_complex* src = (const gr_complex*)raw.data();

  gr::doa::autocorrelate::sptr ac = gr::doa::autocorrelate::make(M, N, overlap, avg);
  /* optional: the antenna_correction block's config file folded into the covariance block */
  if (std::getenv("DOA_HARNESS_ANTENNA_CFG")) ac->set_antenna_config(std::getenv("DOA_HARNESS_ANTENNA_CFG"));
  gr::doa::MUSIC_lin_array::sptr mus = gr::doa::MUSIC_lin_array::make(d, T, M, P);
  gr::doa::find_local_max::sptr flm = gr::doa::find_local_max::make(K, P, 0.0f, 180.0f);
  gr::doa::rootMUSIC_linear_array::sptr rm = gr::doa::rootMUSIC_linear_array::make(d, T, M);
  /* the fused block: same inputs as autocorrelate, same outputs as find_local_max */
  gr::doa::music_chain::sptr mc = gr::doa::music_chain::make(M, N, overlap, avg, d, T, P, K, 0.0f, 180.0f);
  if (std::getenv("DOA_HARNESS_ANTENNA_CFG")) mc->set_antenna_config(std::getenv("DOA_HARNESS_ANTENNA_CFG"));
  if (mc->input_signature()->min_streams() != M || mc->output_signature()->max_streams() != 2 || (int)mc->history() != overlap + 1) {
    std::fprintf(stderr, "music_chain io signature mismatch\n");
    return 3;
  }

  /* autocorrelate + rootMUSIC_linear_array in one block: autocorrelate's inputs, rootMUSIC's port 0 */
  gr::doa::rootmusic_chain::sptr rc = gr::doa::rootmusic_chain::make(M, N, overlap, avg, d, T);
  if (std::getenv("DOA_HARNESS_ANTENNA_CFG")) rc->set_antenna_config(std::getenv("DOA_HARNESS_ANTENNA_CFG"));
  if (rc->input_signature()->min_streams() != M || rc->output_signature()->sizeof_stream_item(0) != (int)sizeof(float) * T ||
      (int)rc->history() != overlap + 1) {
    std::fprintf(stderr, "rootmusic_chain io signature mismatch\n");
    return 3;
  }
  /* optional: the sc16-fed fused block next to it */
  gr::doa::music_chain::sptr mc16;
  std::vector<char> raw16;
  if (std::getenv("DOA_HARNESS_SC16_IN")) {
    raw16 = slurp(std::getenv("DOA_HARNESS_SC16_IN"));
    if (raw16.size() != L * M * 4) { std::fprintf(stderr, "sc16 input length mismatch\n"); return 2; }
    mc16 = gr::doa::music_chain::make_sc16(M, N, overlap, avg, d, T, P, K, 0.0f, 180.0f, 1.0f / 32768);
    if (mc16->input_signature()->sizeof_stream_item(0) != 4 || (int)mc16->history() != overlap + 1) {
      std::fprintf(stderr, "music_chain (sc16) io signature mismatch\n");
      return 3;
    }
  }

  /* io signatures are the reference's (lib/autocorrelate_impl.cc:48-50 etc.) */
  if (ac->input_signature()->min_streams() != M || ac->output_signature()->sizeof_stream_item(0) != (int)sizeof(gr_complex) * M * M ||
      mus->output_signature()->sizeof_stream_item(0) != (int)sizeof(float) * P || flm->output_signature()->max_streams() != 2 ||
      rm->output_signature()->max_streams() != T || (int)ac->history() != overlap + 1) {
    std::fprintf(stderr, "io signature mismatch\n");
    return 3;
  }

  /* scheduler hints of the fused chain blocks: a real scheduler would only offer them multiples of
   * output_multiple(); this harness keeps offering arbitrary n to every block (the blocks accept any n) and checks the hints */
  if (mc->output_multiple() < 2 || mc->min_output_buffer() < 4L * mc->output_multiple() || rc->output_multiple() != mc->output_multiple() ||
      ac->output_multiple() != 1) {
    std::fprintf(stderr, "scheduler hints: music_chain multiple %d buffer %ld, rootmusic_chain multiple %d\n", mc->output_multiple(),
                 mc->min_output_buffer(), rc->output_multiple());
    return 5;
  }
  const int hop = N - overlap;
  /* GNU Radio pre-fills history()-1 zeros in front of the stream; gr-doa's QA vectors are laid out so that the first
   * snapshot starts at sample 0, i.e. the scheduler view is: read pointer at sample 0, `overlap` samples of look-ahead
   * required beyond hop*n.  Emulate exactly that: available = L, a call may produce n frames iff hop*n + overlap <= avail. */
  std::vector<gr_complex> Rbuf; std::vector<float> spec, val, loc, aoa, cval, cloc, sval, sloc, caoa;
  size_t rd = 0;                       /* read pointer (samples) shared by all channels */
  unsigned lcg = 12345;
  size_t frames_total = 0;
  for (;;) {
    const size_t avail = L - rd;
    int max_n = avail >= (size_t)overlap ? (int)((avail - overlap) / hop) : 0;
    if (max_n <= 0) break;
    lcg = lcg * 1664525u + 1013904223u;
    int n = 1 + (int)((lcg >> 16) % 97);                 /* scheduler-chosen batch: 1..97 items */
    n = std::min(n, max_n);
    gr_vector_int need(M, 0);
    ac->forecast(n, need);
    if ((size_t)need[0] + overlap > avail) { std::fprintf(stderr, "forecast asks for more than available\n"); return 3; }
    gr_vector_int nin(M, (int)avail);
    gr_vector_const_void_star ins(M);
    for (int k = 0; k < M; ++k) ins[k] = src + (size_t)k * L + rd;
    Rbuf.resize((frames_total + n) * (size_t)M * M);
    gr_vector_void_star outs(1, Rbuf.data() + frames_total * (size_t)M * M);
    const int produced = ac->general_work(n, nin, ins, outs);
    if (produced != n || ac->last_consumed() != hop * n) { std::fprintf(stderr, "autocorrelate produced %d consumed %d\n", produced, ac->last_consumed()); return 3; }
    {   /* the fused block sees the same scheduler call */
      cval.resize((frames_total + n) * (size_t)K); cloc.resize(cval.size());
      gr_vector_int need2(M, 0);
      mc->forecast(n, need2);
      if (need2[0] != need[0]) { std::fprintf(stderr, "music_chain forecast differs from autocorrelate\n"); return 3; }
      gr_vector_void_star out_c(2); out_c[0] = cval.data() + frames_total * (size_t)K; out_c[1] = cloc.data() + frames_total * (size_t)K;
      if (mc->general_work(n, nin, ins, out_c) != n || mc->last_consumed() != hop * n) { std::fprintf(stderr, "music_chain produced/consumed mismatch\n"); return 3; }
    }
    {   /* and so does the Root-MUSIC chain block */
      caoa.resize((frames_total + n) * (size_t)T);
      gr_vector_void_star out_r(1, caoa.data() + frames_total * (size_t)T);
      if (rc->general_work(n, nin, ins, out_r) != n || rc->last_consumed() != hop * n) { std::fprintf(stderr, "rootmusic_chain produced/consumed mismatch\n"); return 3; }
    }
    if (mc16) {
      sval.resize((frames_total + n) * (size_t)K); sloc.resize(sval.size());
      gr_vector_const_void_star ins16(M);
      for (int k = 0; k < M; ++k) ins16[k] = raw16.data() + ((size_t)k * L + rd) * 4;
      gr_vector_void_star out_s(2); out_s[0] = sval.data() + frames_total * (size_t)K; out_s[1] = sloc.data() + frames_total * (size_t)K;
      if (mc16->general_work(n, nin, ins16, out_s) != n || mc16->last_consumed() != hop * n) { std::fprintf(stderr, "music_chain (sc16) produced/consumed mismatch\n"); return 3; }
    }
    rd += ac->last_consumed();

    /* downstream sync blocks see the same n items */
    spec.resize((frames_total + n) * (size_t)P); val.resize((frames_total + n) * (size_t)K); loc.resize(val.size());
    aoa.resize((frames_total + n) * (size_t)T);
    gr_vector_int nin1(1, n);
    gr_vector_const_void_star in_R(1, Rbuf.data() + frames_total * (size_t)M * M);
    gr_vector_void_star out_spec(1, spec.data() + frames_total * (size_t)P);
    if (mus->general_work(n, nin1, in_R, out_spec) != n) return 4;
    gr_vector_const_void_star in_spec(1, spec.data() + frames_total * (size_t)P);
    gr_vector_void_star out_pk(2); out_pk[0] = val.data() + frames_total * (size_t)K; out_pk[1] = loc.data() + frames_total * (size_t)K;
    if (flm->general_work(n, nin1, in_spec, out_pk) != n) return 4;
    gr_vector_void_star out_aoa(1, aoa.data() + frames_total * (size_t)T);
    if (rm->general_work(n, nin1, in_R, out_aoa) != n) return 4;
    frames_total += n;
  }
  dump(prefix + ".R.c64", Rbuf.data(), Rbuf.size() * sizeof(gr_complex));
  dump(prefix + ".spec.f32", spec.data(), spec.size() * sizeof(float));
  dump(prefix + ".val.f32", val.data(), val.size() * sizeof(float));
  dump(prefix + ".loc.f32", loc.data(), loc.size() * sizeof(float));
  dump(prefix + ".aoa.f32", aoa.data(), aoa.size() * sizeof(float));
  dump(prefix + ".cval.f32", cval.data(), cval.size() * sizeof(float));
  dump(prefix + ".cloc.f32", cloc.data(), cloc.size() * sizeof(float));
  dump(prefix + ".caoa.f32", caoa.data(), caoa.size() * sizeof(float));
  if (mc16) {
    dump(prefix + ".sval.f32", sval.data(), sval.size() * sizeof(float));
    dump(prefix + ".sloc.f32", sloc.data(), sloc.size() * sizeof(float));
  }
  std::printf("frames %zu\n", frames_total);
  return 0;
}
