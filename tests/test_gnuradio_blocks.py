"""The GNU Radio block sources (gr_doa_b200/gnuradio/lib/*_impl.cc) compiled against the compile-only shim and driven by
the fake scheduler (history / forecast / consume_each, scheduler-chosen noutput_items), like python/qa_*.py drive the
reference blocks through a top_block.  CPU part: they compile and link.  GPU part: outputs against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests import parity
from tests.conftest import ROOT


def harness():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_harness", os.path.join(ROOT, "gr_doa_b200", "gnuradio", "build_harness.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_block_sources_compile_against_the_shim():
    from gr_doa_b200 import build as lib_build
    lib_build.build()
    exe = harness()
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_boundary_files_keep_the_reference_interface():
    base = os.path.join(ROOT, "gr_doa_b200", "gnuradio")
    want = {
        "doa_autocorrelate.xml": "doa.autocorrelate($inputs, $snapshot_size, $overlap_size, $avg_method)",
        "doa_MUSIC_lin_array.xml": "doa.MUSIC_lin_array($norm_spacing, $num_targets, $inputs, $pspectrum_len)",
        "doa_rootMUSIC_linear_array.xml": "doa.rootMUSIC_linear_array($norm_spacing, $num_targets, $inputs)",
        "doa_find_local_max.xml": "doa.find_local_max($num_max_vals, $vector_len, $x_min, $x_max)",
        "doa_calibrate_lin_array.xml": "doa.calibrate_lin_array($norm_spacing, $num_ant_ele, $pilot_angle)",
    }
    for fn, make in want.items():
        assert "<make>" + make + "</make>" in open(os.path.join(base, "grc", fn)).read()
    swig = open(os.path.join(base, "swig", "doa_swig.i")).read()
    for blk in ("autocorrelate", "MUSIC_lin_array", "rootMUSIC_linear_array", "find_local_max", "calibrate_lin_array"):
        assert f"GR_SWIG_BLOCK_MAGIC2(doa, {blk});" in swig
        hdr = open(os.path.join(base, "include", "doa", blk + ".h")).read()
        assert "static sptr make(" in hdr and "boost::shared_ptr<" + blk + ">" in hdr


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,overlap,avg,T,P,K", [(4, 2048, 512, 1, 1, 2048, 1), (8, 256, 32, 0, 2, 1024, 2)])
def test_fake_scheduler_flowgraph_matches_oracle(oracle, tmp_path, M, N, overlap, avg, T, P, K):
    from gr_doa_b200 import synth
    exe = harness()
    nframes = 300
    thetas = [60.0] if T == 1 else [50.0, 110.0]
    x = synth.stream_numpy(nframes, M, N, overlap, thetas, seed=77 + M)
    inp = tmp_path / "in.c64"
    x.astype(np.complex64).tofile(inp)
    r = subprocess.run([exe, str(inp), str(M), str(N), str(overlap), str(avg), "0.5", str(T), str(P), str(K), str(tmp_path / "out")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"frames {nframes}" in r.stdout and "Total output items produced: %d" % nframes in r.stdout
    R = np.fromfile(tmp_path / "out.R.c64", np.complex64).reshape(nframes, M * M)
    spec = np.fromfile(tmp_path / "out.spec.f32", np.float32).reshape(nframes, P)
    val = np.fromfile(tmp_path / "out.val.f32", np.float32).reshape(nframes, K)
    loc = np.fromfile(tmp_path / "out.loc.f32", np.float32).reshape(nframes, K)
    aoa = np.fromfile(tmp_path / "out.aoa.f32", np.float32).reshape(nframes, T)
    R_o = oracle.autocorrelate(x, N, overlap, avg)
    assert parity.rel_fro(R, R_o) <= parity.COV_REL_FRO
    q64 = oracle.music_f64(R, 0.5, T, M, P)
    assert parity.spectrum_db_error(spec, oracle.music(R, 0.5, T, M, P), q64) <= parity.SPECTRUM_DB
    v_o, l_o, b_o = oracle.find_local_max(spec, K, 0.0, 180.0)
    assert np.array_equal(val, v_o) and np.array_equal(loc, l_o)          # block-to-block: bit-exact on the GPU spectrum
    a64, d64 = oracle.rootmusic_f64(R, 0.5, T, M, return_dist=True)
    worst, near = parity.root_angles_ok(aoa, a64, d64)
    assert worst <= parity.ROOT_DEG and near <= 6
    assert np.abs(np.sort(loc, 1) - np.sort(np.array(thetas))[None, :]).max() < 2.0     # the reference QA's own bound
    # doa.rootmusic_chain (autocorrelate + rootMUSIC_linear_array in one block): same kernels on the same covariance = same bits
    caoa = np.fromfile(tmp_path / "out.caoa.f32", np.float32).reshape(nframes, T)
    assert np.array_equal(caoa, aoa)
    # the fused block (doa.music_chain: same inputs as autocorrelate, same outputs as find_local_max) under the same scheduler
    cval = np.fromfile(tmp_path / "out.cval.f32", np.float32).reshape(nframes, K)
    cloc = np.fromfile(tmp_path / "out.cloc.f32", np.float32).reshape(nframes, K)
    step = 180.0 / P
    same = np.abs(cloc - loc).max(1) <= 1e-6
    assert same.mean() >= 0.98 and np.abs(cloc - loc).max() <= 1.001 * step      # +-1-bin near-ties only
    assert np.abs(np.sort(cval, 1) - np.sort(val, 1))[same].max() <= 0.05        # dB; the chain refines with v^H G v


@pytest.mark.gpu
def test_fake_scheduler_with_antenna_config(oracle, tmp_path):
    """The autocorrelate block with the antenna_correction config file folded in (set_antenna_config) produces the covariance
    of the corrected streams: the reference's antenna_correction -> autocorrelate pair, one block and one pass fewer."""
    from gr_doa_b200 import synth
    from tests.test_channel_gains import reference_gains
    exe = harness()
    M, N, overlap, avg, T, P, K, n = 4, 256, 64, 1, 1, 512, 1, 23
    gain = [1.0, 0.8, 1.3, 0.6]; phase = [0.0, 0.4, -0.9, 2.2]
    cfg = tmp_path / "antenna.cfg"
    cfg.write_text("".join(f"{g} {p}\n" for g, p in zip(gain, phase)))
    x = synth.stream_numpy(n, M, N, overlap, [60.0], seed=99)
    inp = tmp_path / "in.c64"
    x.astype(np.complex64).tofile(inp)
    env = dict(os.environ, DOA_HARNESS_ANTENNA_CFG=str(cfg))
    r = subprocess.run([exe, str(inp), str(M), str(N), str(overlap), str(avg), "0.5", str(T), str(P), str(K), str(tmp_path / "out")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    R = np.fromfile(str(tmp_path / "out") + ".R.c64", dtype=np.complex64).reshape(-1, M * M)
    g = reference_gains(gain, phase)
    exp = oracle.autocorrelate((g[:, None] * x).astype(np.complex64), N, overlap, avg)
    assert R.shape == exp.shape
    assert parity.rel_fro(R, exp) <= parity.COV_REL_FRO


@pytest.mark.gpu
def test_fake_scheduler_music_chain_fed_sc16_items(tmp_path):
    """doa.music_chain made with make_sc16 (4-byte Complex Int16 items, what UHD delivers before its host-side conversion to
    gr_complex) under the same scheduler calls as the fc32 block fed the converted samples: identical peaks, bit for bit."""
    from gr_doa_b200 import synth
    from tests.test_sc16_input import S15, quantise, to_fc32
    exe = harness()
    M, N, overlap, avg, T, P, K, n = 4, 2048, 512, 1, 1, 2048, 1, 150
    q = quantise(synth.stream_numpy(n, M, N, overlap, [60.0], seed=31))
    inp, inp16 = tmp_path / "in.c64", tmp_path / "in.sc16"
    to_fc32(q, S15).tofile(inp)
    q.tofile(inp16)
    env = dict(os.environ, DOA_HARNESS_SC16_IN=str(inp16))
    r = subprocess.run([exe, str(inp), str(M), str(N), str(overlap), str(avg), "0.5", str(T), str(P), str(K), str(tmp_path / "out")],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert f"frames {n}" in r.stdout
    got = [np.fromfile(str(tmp_path / "out") + ext, np.float32).reshape(n, K) for ext in (".sval.f32", ".sloc.f32")]
    ref = [np.fromfile(str(tmp_path / "out") + ext, np.float32).reshape(n, K) for ext in (".cval.f32", ".cloc.f32")]
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    assert np.abs(got[1] - 60.0).max() < 2.0


@pytest.mark.gpu
def test_fake_scheduler_music_chain_over_a_device_list(tmp_path):
    """DOA_CUDA_DEVICES: one doa.music_chain instance spreads each work() call's frames over the listed devices
    (doa_cuda_multi_run_streams: every device reads its block, overlap included, from the scheduler's buffers).  Same
    scheduler calls, same peaks, bit for bit, as on one device.  (A device listed twice = two independent stream sets on it,
    which is how a one-GPU box exercises the sharding.)"""
    import torch
    from gr_doa_b200 import synth
    exe = harness()
    M, N, overlap, avg, T, P, K, n = 8, 256, 32, 0, 2, 1024, 2, 200
    x = synth.stream_numpy(n, M, N, overlap, [50.0, 110.0], seed=17)
    inp = tmp_path / "in.c64"
    x.astype(np.complex64).tofile(inp)
    outs = {}
    nd = torch.cuda.device_count()
    lists = {"one": "0", "twice": "0,0", "all": ",".join(str(i) for i in range(nd)) + ",0"}
    for tag, devs in lists.items():
        env = dict(os.environ, DOA_CUDA_DEVICES=devs)
        r = subprocess.run([exe, str(inp), str(M), str(N), str(overlap), str(avg), "0.5", str(T), str(P), str(K), str(tmp_path / tag)],
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        outs[tag] = [np.fromfile(str(tmp_path / tag) + ext, np.float32).reshape(n, K) for ext in (".cval.f32", ".cloc.f32")]
    for tag in ("twice", "all"):
        assert np.array_equal(outs[tag][0], outs["one"][0]) and np.array_equal(outs[tag][1], outs["one"][1])
    assert np.abs(np.sort(outs["one"][1], 1) - np.array([50.0, 110.0])[None, :]).max() < 2.0
