// herk_geometry.h -- how the tensor-core HERK (herk_tc.cu) cuts a batch into work units: plain C++, shared by the launcher, the
// kernel's three roles (which must agree on every unit's stage count: a mismatch is a hang) and the CPU test that sweeps it
// (tests/test_herk_geometry.py).
//
// A stage = 16 complex samples of every channel; a frame of N snapshots = spf stages, cut into nseg <= HERK_MAX_SEGS segments of
// seg_len stages (a multiple of the 4-stage accumulator chunk; a function of N alone -- the association of a frame's sum).
// CTA b of `grid` takes the whole frames b, b + grid, ... < nfull; the nframes - nfull TAIL frames (those that do not fill a round of
// the grid) are shared: S CTAs per frame, CTA b the segments [seg0, seg0 + sps) of tail frame b / S.
#pragma once

#ifdef __CUDACC__
#define DOA_HD __host__ __device__
#else
#define DOA_HD
#endif

namespace doa {

constexpr int HERK_CHUNK = 4;        // stages per accumulator chunk
constexpr int HERK_MAX_SEGS = 8;     // segments per frame, also the widest split
constexpr int HERK_WS_FRAMES = 128;  // frames one launch may share (workspace slots)

struct HerkGeometry {
  int spf, seg_len, nseg;            // per frame
  int grid, nfull, S, sps;           // per launch; S = 0: no frame is shared (nfull = nframes)
};

inline HerkGeometry herk_geometry(int nframes, int N, int sms, bool split) {
  HerkGeometry g;
  g.spf = (N + 15) / 16;
  const int per = (g.spf + HERK_MAX_SEGS - 1) / HERK_MAX_SEGS;
  g.seg_len = (per + HERK_CHUNK - 1) / HERK_CHUNK * HERK_CHUNK;
  if (g.seg_len < HERK_CHUNK) g.seg_len = HERK_CHUNK;
  g.nseg = (g.spf + g.seg_len - 1) / g.seg_len;
  g.grid = nframes < sms ? nframes : sms;
  g.nfull = nframes; g.S = 0; g.sps = 0;
  const int t = nframes % sms;
  if (split && t > 0 && t <= HERK_WS_FRAMES && g.nseg >= 2 && sms / t >= 2) {
    const int want = sms / t < g.nseg ? sms / t : g.nseg;
    const int sps = (g.nseg + want - 1) / want;       // segments per CTA
    const int S = (g.nseg + sps - 1) / sps;           // CTAs per shared frame, every one non-empty
    if (S >= 2) { g.S = S; g.sps = sps; g.grid = nframes > sms ? sms : t * S; g.nfull = nframes - t; }
  }
  return g;
}

struct HerkTail { int has, idx, seg0, start, count; };   // CTA b's share of a tail frame (stages [start, start + count) of frame nfull + idx)

DOA_HD inline HerkTail herk_tail(int b, int nframes, int nfull, int S, int sps, int seg_len, int spf) {
  HerkTail t = {0, 0, 0, 0, 0};
  if (b >= (nframes - nfull) * S) return t;
  const int nseg = (spf + seg_len - 1) / seg_len;
  t.has = 1;
  t.idx = b / S;
  t.seg0 = (b % S) * sps;
  t.start = t.seg0 * seg_len;
  int seg1 = t.seg0 + sps; if (seg1 > nseg) seg1 = nseg;
  int end = seg1 * seg_len; if (end > spf) end = spf;
  t.count = end - t.start;
  return t;
}

// whole frames of CTA b
DOA_HD inline int herk_whole_frames(int b, int grid, int nfull) { return nfull > b ? (nfull - 1 - b) / grid + 1 : 0; }

}  // namespace doa
