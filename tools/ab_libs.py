"""A/B of whole-library builds on one box: each argument is a libdoa_cuda*.so; every round times every library in its own
process (cfg3 fused chain, cfg5's three stages), round-robin so that clock/thermal drift hits all alike.
usage: ab_libs.py gr_doa_b200/_ab/libA.so gr_doa_b200/_ab/libB.so ...      (child: ab_libs.py --child <lib>)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(path):
    import torch
    from gr_doa_b200 import _lib, synth
    _lib.LIB_PATH = os.path.abspath(path)
    import gr_doa_b200 as doa

    def t(fn, reps=8):
        for _ in range(3): fn()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        return round(best, 4)
    out = {}
    x, _ = synth.frames_torch(65536, 8, 2048, [40.0, 90.0, 140.0], jitter_deg=2.0, device="cuda", chunk=2048)
    ch = doa.DoaChain(8, 2048, 0, 0, 0.5, 3, 4096, 3, max_frames=65536)
    out["cfg3"] = t(lambda: ch.run_device(x))
    del x, ch
    x, _ = synth.frames_torch(65536, 16, 1024, [40.0, 90.0, 140.0], jitter_deg=2.0, device="cuda", chunk=2048)
    ch = doa.DoaChain(16, 1024, 0, 0, 0.5, 3, 4096, 3, max_frames=65536)
    out["cfg5"] = t(lambda: ch.run_device(x), 4)
    ac = doa.autocorrelate(16, 1024, 0, 0, max_frames=65536)
    out["cov16"] = t(lambda: ac.work_device(x), 4)
    R = ac.work_device(x)
    mu = doa.MUSIC_lin_array(0.5, 3, 16, 4096, max_frames=65536)
    out["eig16"] = t(lambda: mu.noise_subspace_device(R), 4)
    del x, ch
    x, _ = synth.frames_torch(262144, 4, 2048, [60.0], jitter_deg=2.0, device="cuda", chunk=4096)
    ch = doa.DoaChain(4, 2048, 0, 0, 0.5, 1, 2048, 1, max_frames=262144)
    out["cfg1"] = t(lambda: ch.run_device(x), 4)
    del x, ch
    torch.cuda.empty_cache()
    x, _ = synth.frames_torch(512, 64, 16384, [30.0 + 120.0 * i / 7 for i in range(8)], jitter_deg=2.0, device="cuda", chunk=32)
    ac = doa.autocorrelate(64, 16384, 0, 0, max_frames=512)
    out["cov64"] = t(lambda: ac.work_device(x), 4)
    ch = doa.DoaChain(64, 16384, 0, 0, 0.5, 8, 16384, 8, max_frames=512)
    out["cfg4"] = t(lambda: ch.run_device(x), 4)
    out["cov64_checksum"] = float(torch.view_as_real(ac.work_device(x)).double().abs().sum().item())
    print(json.dumps(out))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        libs, res = sys.argv[1:], {}
        for rnd in range(3):
            for l in libs:
                r = subprocess.run([sys.executable, __file__, "--child", l], capture_output=True, text=True)
                try:
                    d = json.loads(r.stdout.strip().splitlines()[-1])
                except Exception:
                    print(l, "failed:", r.stderr[-400:]); continue
                for k, v in d.items():
                    res.setdefault(l, {}).setdefault(k, []).append(v)
        for l in libs:
            print(os.path.basename(l), {k: (min(v), sorted(v)[len(v) // 2]) for k, v in res.get(l, {}).items()}, "(min, median of rounds; ms)", flush=True)
