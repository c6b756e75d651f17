"""Development check run on a GPU box: each stage against the oracle, plus quick timings. Not a test."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa

def relfro(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def stage_checks(M, T, N, P, K, thetas, B=512, avg=0, d=0.5):
    print(f"--- M={M} T={T} N={N} P={P} K={K} avg={avg}")
    fr, th = synth.frames_numpy(B, M, N, thetas, jitter_deg=3, snr_db=10, seed=11 + M)
    R_o = O.autocorrelate_frames(fr, avg, nthreads=8)
    ac = doa.autocorrelate(M, N, 0, avg, max_frames=B)
    R_g = ac.work_device(torch.from_numpy(fr).cuda()).cpu().numpy()
    print("cov relfro", relfro(R_g, R_o))
    mus = doa.MUSIC_lin_array(d, T, M, P, max_frames=B)
    S_g = mus.work(R_o)
    S_o = O.music(R_o, d, T, M, P, nthreads=8)
    Q64 = O.music_f64(R_o, d, T, M, P, nthreads=8)
    mask = Q64 > 1e-2 * Q64.max(1, keepdims=True)
    dd = (S_g - S_o); dd = dd - np.median(dd, axis=1, keepdims=True)
    print("spectrum: max |dB diff| (offset removed, outside nulls)", np.abs(dd[mask]).max(), "raw max", np.abs(S_g - S_o).max())
    loc_t, th_t, V_t = mus.tables(); lo, tho, Vo = O.music_tables(d, M, P)
    print("tables bit-equal:", np.array_equal(loc_t, lo), np.array_equal(th_t, tho), np.array_equal(V_t.view(np.float32), Vo.view(np.float32)))
    flm = doa.find_local_max(K, P, 0.0, 180.0, max_frames=B)
    v_g, l_g, b_g = flm.work(S_o, return_bins=True)
    v_o, l_o, b_o = O.find_local_max(S_o, K, 0.0, 180.0, nthreads=8)
    print("find_local_max bit-exact:", np.array_equal(v_g, v_o), np.array_equal(l_g, l_o), np.array_equal(b_g, b_o))
    ch = doa.DoaChain(M, N, 0, avg, d, T, P, K, max_frames=B)
    cv, cl, cb = ch.run_host(fr)
    mism = (cb != b_o).any(1)
    print("chain bins mismatch frames:", mism.sum(), "/", B, " max|val diff|", np.abs(cv - v_o)[~mism].max() if (~mism).any() else None,
          " loc equal on matching:", np.array_equal(cl[~mism], l_o[~mism]))
    if mism.any():
        i = np.where(mism)[0][0]; print("  e.g.", cb[i], b_o[i], cv[i], v_o[i])
    rm = doa.rootMUSIC_linear_array(d, T, M, max_frames=B)
    a_g = rm.work(R_o)
    a64 = O.rootmusic_f64(R_o, d, T, M, nthreads=8); a32 = O.rootmusic(R_o, d, T, M, nthreads=8)
    print("rootmusic max |gpu-f64| deg", np.nanmax(np.abs(a_g - a64)), " |oracle32-f64|", np.nanmax(np.abs(a32 - a64)), "nan", np.isnan(a_g).sum())

def timing(B=65536, M=8, N=2048, T=3, P=4096, K=3):
    print(f"--- timing B={B} M={M} N={N} P={P}")
    x, _ = synth.frames_torch(B, M, N, [40, 90, 140] if T == 3 else [60], jitter_deg=5, device="cuda")
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    ch.set_profiling(True)
    L = _lib.lib()
    for variant in (2, 1):
        doa.set_default_option("cov_groups", variant)
        for it in range(4):
            torch.cuda.synchronize(); t = time.time()
            out = ch.run_device(x); torch.cuda.synchronize(); dt = time.time() - t
            ms = ch.stage_ms()
        gb = B * M * N * 8 / 1e9
        print(f"variant {variant}: total {dt*1e3:.3f} ms  stages(cov,eig,scan) = {ms[0]:.3f} {ms[1]:.3f} {ms[2]:.3f} ms  cov GB/s {gb/ms[0]*1e3:.0f}  chain GB/s {gb/sum(ms)*1e3:.0f}")
    # sanity vs oracle on a few frames
    sub = x[:64].cpu().numpy()
    v_o, l_o, b_o = O.chain_frames(sub, 0, 0.5, T, P, K, nthreads=8)
    print("bins equal on 64 frames:", (out[2][:64].cpu().numpy() == b_o).all(1).mean())

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    stage_checks(8, 3, 2048, 4096, 3, [40, 90, 140])
    stage_checks(4, 1, 2048, 2048, 1, [60], avg=1)
    stage_checks(4, 2, 2048, 1024, 2, [50, 110], avg=1)
    stage_checks(16, 3, 1024, 4096, 3, [40, 90, 140], B=256)
    stage_checks(6, 2, 500, 1000, 2, [50, 110], B=128)
    timing()
    os.system("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv")
