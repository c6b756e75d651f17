"""Shared parity criteria.

Tolerances come from BASELINE.json's north_star: covariance rel. Frobenius <= 1e-5; pseudo-spectra within 1e-3 dB
outside deep nulls (compared with the common max-offset removed, because the reference normalises by 1/Q at its deepest
null, the single worst-conditioned number of the whole path); peak bins identical except documented near-ties;
Root-MUSIC within 1e-4 degree (against the float64 twin: the reference's own float32 cgeev is ~1e-2 degree noisy).

A NEAR-TIE is a bin disagreement of one bin whose two candidate bins have float64 null-spectrum values closer than the
float32 reference's own distance from that float64 spectrum on the frame (times a small safety factor): the reference's
arithmetic itself cannot tell the two bins apart.
"""
import numpy as np

COV_REL_FRO = 1e-5
SPECTRUM_DB = 1e-3
ROOT_DEG = 1e-4
NEAR_TIE_SAFETY = 8.0


ROOT_NEAR_CIRCLE = 4e-4      # ~ sqrt(float32 coefficient noise): closer than this, inside/outside is not decidable


def root_angles_ok(aoa_gpu, aoa64, dist64):
    """Root-MUSIC criterion: within ROOT_DEG of the float64 twin on every well-conditioned frame.  A frame whose selected
    root sits within ROOT_NEAR_CIRCLE of the unit circle has a nearly double root (z, 1/conj z): coefficient noise of 1e-7
    moves such roots by ~sqrt(1e-7) and decides which twin is "strictly inside" (lib/rootMUSIC_linear_array_impl.cc:125);
    the reference's own float32 cgeev flips on those frames too.  Returns (worst error on good frames, #near-circle)."""
    good = np.nanmin(dist64, axis=1) >= ROOT_NEAR_CIRCLE
    worst = float(np.abs(aoa_gpu[good] - aoa64[good]).max()) if good.any() else 0.0
    return worst, int((~good).sum())


def rel_fro(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def spectrum_db_error(spec_gpu, spec_ref, q64):
    """max |dB difference| outside deep nulls (Q > 1e-2 max Q per frame), per-frame common offset removed."""
    mask = q64 > 1e-2 * q64.max(axis=1, keepdims=True)
    diff = spec_gpu.astype(np.float64) - spec_ref.astype(np.float64)
    worst = 0.0
    for f in range(diff.shape[0]):
        d = diff[f][mask[f]]
        d = d[np.isfinite(d)]
        worst = max(worst, float(np.abs(d - np.median(d)).max()))
    return worst


def classify_bins(bins_gpu, bins_ref, q64, q32_ref):
    """Returns (frames_with_any_difference, frames_not_explained_as_near_ties)."""
    bg, br = np.sort(bins_gpu, axis=1), np.sort(bins_ref, axis=1)
    diff_frames = np.where((bg != br).any(axis=1))[0]
    unexplained = []
    for f in diff_frames:
        noise = float(np.abs(q32_ref[f].astype(np.float64) - q64[f]).max())
        ok = True
        for a, b in zip(bg[f], br[f]):
            if a == b:
                continue
            if abs(int(a) - int(b)) > 1 or abs(q64[f, a] - q64[f, b]) > NEAR_TIE_SAFETY * noise:
                ok = False
        if not ok:
            unexplained.append(int(f))
    return len(diff_frames), unexplained


def peak_value_bound_db(q64, bins_ref, q32_ref):
    """Per-entry bound on |peak height difference| in dB: heights are 10*log10(Qmin/Q_k) with both Q's carrying the
    reference's own absolute float32 noise tau; first-order propagation with the same safety factor."""
    tau = np.abs(q32_ref.astype(np.float64) - q64).max(axis=1, keepdims=True) * NEAR_TIE_SAFETY
    qk = np.take_along_axis(q64, bins_ref.astype(np.int64), axis=1)
    qmin = q64.min(axis=1, keepdims=True)
    return (10.0 / np.log(10.0)) * tau * (1.0 / qk + 1.0 / qmin) + 1e-4
