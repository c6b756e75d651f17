"""Independent float64 numpy restatement of the reference's OCTAVE model (examples/@wpi_twinrx_doa_testbench/*.m and
python/test00{1,2}_findpeaks.m), used only to pin the C++ oracle.  The reference's QA gets its expected values from
this model through a live Octave; here it is restated so the same known-answer cases run without Octave."""
import numpy as np


def octave_autocorrelate(xx, len_ss, overlap_size, FB):
    """autocorrelate.m:23-48.  xx: [len_input][num_inputs].  Returns S_x as the flattened column-major stream."""
    hop = len_ss - overlap_size
    num_ss = (xx.shape[0] - overlap_size) // hop
    M = xx.shape[1]
    J = np.fliplr(np.eye(M))
    out = []
    for ii in range(num_ss):
        x = xx[ii * hop: ii * hop + len_ss, :].astype(np.complex128)
        S = x.T @ np.conj(x) / len_ss                                   # transpose(x)*conj(x)/len_ss  (:38)
        if FB:
            S = 0.5 * S + 0.5 * J @ np.conj(S) @ J / len_ss             # (:41-45) note the extra /len_ss
        out.append(S.flatten(order="F"))
    return np.array(out)


def octave_amv(theta_rad, M, d):
    loc = d * np.arange((M - 1) / 2, -(M - 1) / 2 - 1, -1)              # wpi_twinrx_doa_testbench.m:60-61
    return np.exp(-1j * 2 * np.pi * np.cos(theta_rad) * loc)             # :64


def octave_music_input(num_ss, len_ss, overlap_size, M, d, doas_deg, FB, rng, snr_db=1000.0):
    """music_test_input_gen.m:14-118 without perturbation: tones at w = pi/denom, noise at `snr` dB."""
    hop = len_ss - overlap_size
    L = hop * num_ss + overlap_size
    D = np.deg2rad(np.asarray(doas_deg, dtype=np.float64))
    denom = rng.permutation(M) + 1
    w = np.pi / denom[: len(D)]
    V = np.stack([octave_amv(t, M, d) for t in D], axis=1)              # [M][T]
    xx = (V @ np.exp(1j * w[:, None] * np.arange(1, L + 1)[None, :])).T  # [L][M]
    snr_lin = 10.0 ** (snr_db / 10.0)
    En = (np.abs(xx) ** 2).sum(0) / L
    sig = np.sqrt(En / snr_lin / 2.0)
    xx = xx + (rng.standard_normal((L, M)) + 1j * rng.standard_normal((L, M))) * sig[None, :]
    return xx


def octave_music_doa(S_flat, M, T, d, P):
    """MUSIC.m:21-50 -> arg-max angle (degrees) per snapshot, float64."""
    theta = np.arange(P) * 180.0 / P
    A = np.stack([octave_amv(np.deg2rad(t), M, d) for t in theta], axis=1)   # [M][P]
    out = []
    for s in S_flat:
        S = s.reshape(M, M, order="F")
        w, E = np.linalg.eigh(S)
        Un = E[:, : M - T]
        Q = 1.0 / np.real(np.einsum("mp,mn,np->p", A.conj(), Un @ Un.conj().T, A))
        out.append(theta[np.argmax(Q)])
    return np.array(out)


def octave_rmusic(S_flat, M, T, d):
    """rMUSIC.m:20-60, float64."""
    out = []
    for s in S_flat:
        S = s.reshape(M, M, order="F")
        w, E = np.linalg.eigh(S)
        Un = E[:, : M - T]
        G = Un @ Un.conj().T
        u = np.array([np.trace(G, offset=l) for l in range(-(M - 1), M)])    # u(l+N) = sum(diag(U_N_sq, l))
        u = u[::-1]
        r = np.roots(u / u[0])
        dist = 1 - np.abs(r)
        keep = dist >= 0
        r, dist = r[keep], dist[keep]
        sel = r[np.argsort(dist)[:T]]
        out.append(np.sort(np.arccos(np.angle(sel) / (2 * np.pi * d)) * 180 / np.pi))
    return np.array(out)


def findpeaks_test_vector(which, vector_len):
    """python/test001_findpeaks.m:5-8 and test002_findpeaks.m:5-8."""
    t = 2 * np.pi * np.linspace(0, 1, vector_len)
    if which == 1:
        y = np.sin(3.14 * t) + 0.5 * np.cos(6.09 * t) + 0.1 * np.sin(10.11 * t + 1 / 6) + 0.1 * np.sin(15.3 * t + 1 / 3)
    else:
        y = np.sin(0.25 * 3.14 * t) + 5 * np.sin(6.09 * t) + 0.6 * np.cos(1.11 * t + 1 / 6) + 2 * np.sin(5.3 * t + 1 / 3)
    return np.abs(y), t


def octave_findpeaks_topk(data, K):
    """findpeaks() on smooth positive data = strict interior local maxima; then sort descending, keep K (test00x:9-13)."""
    d = np.asarray(data, dtype=np.float64)
    idx = np.where((d[1:-1] > d[:-2]) & (d[1:-1] > d[2:]))[0] + 1
    order = np.argsort(-d[idx], kind="stable")[:K]
    return d[idx][order], idx[order]
