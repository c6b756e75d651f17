"""CPU: the work-unit geometry of the tensor-core HERK (gr_doa_b200/csrc/herk_geometry.h), swept by a small C++ program.  The kernel's
MMA warp, converters and adders each derive their CTA's units from these functions; a disagreement or an empty unit is a hang on the
GPU, a gap or an overlap a wrong covariance -- so the invariants are checked here, over every batch size and many snapshot sizes."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECKER = r'''
#include <cstdio>
#include <vector>
#include "herk_geometry.h"
using namespace doa;
int main() {
  const int Ns[] = {2, 16, 30, 64, 66, 256, 528, 1000, 1024, 2048, 4096, 5000, 16384, 16386, 65536};
  const int SMs[] = {148, 132, 108, 4, 1};
  long long cases = 0, shared = 0;
  for (int sms : SMs) for (int N : Ns) for (int split = 0; split < 2; ++split) for (int nframes = 1; nframes <= 3 * sms + 5 && nframes <= 700; ++nframes) {
    const HerkGeometry g = herk_geometry(nframes, N, sms, split != 0);
    ++cases;
    if (g.seg_len % HERK_CHUNK || g.seg_len < HERK_CHUNK || g.nseg < 1 || g.nseg > HERK_MAX_SEGS || (g.nseg - 1) * g.seg_len >= g.spf || g.nseg * g.seg_len < g.spf) { printf("segments N=%d\n", N); return 1; }
    // the segmentation is a function of N alone
    const HerkGeometry g1 = herk_geometry(1, N, sms, false);
    if (g1.seg_len != g.seg_len || g1.nseg != g.nseg) { printf("segmentation depends on the batch N=%d\n", N); return 1; }
    if (g.grid < 1 || g.grid > sms || g.nfull < 0 || g.nfull > nframes) { printf("grid N=%d B=%d\n", N, nframes); return 1; }
    if (!split && g.S != 0) return 2;
    if (g.S == 0 && g.nfull != nframes) return 3;
    const int ntail = nframes - g.nfull;
    if (g.S > 0) { ++shared; if (g.S < 2 || g.S > HERK_MAX_SEGS || ntail < 1 || ntail > HERK_WS_FRAMES || ntail * g.S > g.grid || g.nfull % g.grid != 0) { printf("split N=%d B=%d\n", N, nframes); return 4; } }
    // every stage of every frame is taken exactly once; no unit is empty; tail units are whole segments
    std::vector<std::vector<int>> cover(nframes, std::vector<int>(g.spf, 0));
    std::vector<int> tickets(ntail > 0 ? ntail : 1, 0);
    for (int b = 0; b < g.grid; ++b) {
      const int whole = herk_whole_frames(b, g.grid, g.nfull);
      for (int u = 0; u < whole; ++u) { const long long f = b + (long long)u * g.grid; if (f >= g.nfull) return 5; for (int s = 0; s < g.spf; ++s) ++cover[f][s]; }
      const HerkTail t = herk_tail(b, nframes, g.nfull, g.S, g.sps, g.seg_len, g.spf);
      if (t.has) {
        if (g.S == 0 || t.idx < 0 || t.idx >= ntail || t.count <= 0 || t.start % g.seg_len || t.seg0 * g.seg_len != t.start) { printf("tail unit N=%d B=%d b=%d\n", N, nframes, b); return 6; }
        if ((t.start + t.count) % g.seg_len && t.start + t.count != g.spf) return 7;
        if (t.seg0 + (t.count + g.seg_len - 1) / g.seg_len > g.nseg) return 8;
        for (int s = t.start; s < t.start + t.count; ++s) ++cover[g.nfull + t.idx][s];
        ++tickets[t.idx];
      } else if (whole == 0) { printf("idle CTA N=%d B=%d b=%d\n", N, nframes, b); return 9; }
    }
    for (int f = 0; f < nframes; ++f) for (int s = 0; s < g.spf; ++s) if (cover[f][s] != 1) { printf("coverage N=%d B=%d frame %d stage %d: %d\n", N, nframes, f, s, cover[f][s]); return 10; }
    for (int i = 0; i < ntail; ++i) if (tickets[i] != g.S) { printf("tickets N=%d B=%d\n", N, nframes); return 11; }   // the fold waits for exactly S tickets
  }
  printf("ok %lld cases, %lld with shared frames\n", cases, shared);
  return 0;
}
'''


def test_herk_work_units_cover_every_stage_once(tmp_path):
    src = tmp_path / "check.cpp"
    src.write_text(CHECKER)
    exe = tmp_path / "check"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "gr_doa_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok")
    ncases, nshared = int(out.stdout.split()[1]), int(out.stdout.split()[3])
    assert ncases > 10000 and nshared > 1000
