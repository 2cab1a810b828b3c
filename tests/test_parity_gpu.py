"""GPU parity tests: the CUDA path, called through the C ABI, against (a) the golden vectors
the unmodified reference produced and (b) the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest

from conftest import GOLD, assert_columns_close, rel_to_frame_scale
from opensmile_b200 import Plan, components_mfcc12_0_d_a, pack_utterances
from opensmile_b200.synth import voiced_pcm
from oracle import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5   # relative to the per-frame vector scale, float32 (north_star)


@pytest.fixture(scope="module")
def plan16():
    p = Plan(components_mfcc12_0_d_a(16000.0), "lld", device=0)
    yield p
    p.close()


@pytest.fixture(scope="module")
def plan44():
    p = Plan(components_mfcc12_0_d_a(44100.0), "lld", device=0)
    yield p
    p.close()


def test_golden_example_wav_config0(plan44):
    g = np.load(os.path.join(GOLD, "mfcc_example_44k1.npz"))
    out = plan44.run_host(g["pcm"], np.array([0, g["pcm"].size], np.int64))
    assert out.shape == (202, 39)                       # bit-exact frame count
    assert rel_to_frame_scale(out, g["lld"]) < TOL
    assert_columns_close(out, g["lld"])                 # every column against its own scale


def test_golden_synth16k(plan16):
    g = np.load(os.path.join(GOLD, "mfcc_synth16k_s0.npz"))
    pcm = voiced_pcm(80000, 16000, seed=0)
    out = plan16.run_host(pcm, np.array([0, pcm.size], np.int64))
    assert out.shape == (498, 39)
    assert rel_to_frame_scale(out, g["lld"]) < TOL
    assert_columns_close(out, g["lld"])
    # the regression stages are exact float arithmetic on the statics: recomputing them on the
    # CPU from the GPU's own statics must match the GPU's delta columns bit for bit
    d = oracle.delta(out[:, :13], 2)
    dd = oracle.delta(d, 2)
    assert np.array_equal(d[:498], out[:, 13:26])
    assert np.array_equal(dd[:498], out[:, 26:39])


@pytest.mark.parametrize("lens", [
    [16000, 8123, 400, 399, 0, 561, 5000],       # ragged, too-short and empty utterances
    [80240] * 5,                                  # exactly 500 frames each (BASELINE cfg 2 shape)
    [400 + 160 * 31, 400 + 160 * 32, 400 + 160 * 33],   # tile boundaries (32 frames per tile)
])
def test_batch_vs_oracle_16k(plan16, lens):
    utts = [voiced_pcm(n, 16000, seed=100 + i) for i, n in enumerate(lens)]
    pcm, off = pack_utterances(utts)
    out = plan16.run_host(pcm, off)
    fo = plan16.frame_offsets(off)
    assert out.shape[0] == fo[-1]
    for u, x in enumerate(utts):
        ref = oracle.mfcc_d_a(x, 16000.0)
        got = out[fo[u]:fo[u + 1]]
        assert got.shape == ref.shape
        if ref.shape[0]:
            assert rel_to_frame_scale(got, ref) < TOL


def test_batch_vs_oracle_44k1_stereo():
    comps = components_mfcc12_0_d_a(44100.0, n_channels=2)
    p = Plan(comps, "lld", device=0)
    utts = [voiced_pcm(n, 44100, seed=200 + i, n_chan=2) for i, n in enumerate([30000, 1103, 20011])]
    pcm, off = pack_utterances(utts, n_chan=2)
    out = p.run_host(pcm, off)
    fo = p.frame_offsets(off)
    for u, x in enumerate(utts):
        ref = oracle.mfcc_d_a(x, 44100.0, n_chan=2)
        got = out[fo[u]:fo[u + 1]]
        assert got.shape == ref.shape
        assert rel_to_frame_scale(got, ref) < TOL
    p.close()


def test_batch_vs_oracle_48k_stereo_narrow_tiles():
    """48 kHz stereo: 1200-sample frames (FFT 2048) with a 480-sample stride do not fit the full
    16-frame tile into shared memory -> the half-width (8-frame) kernel instance runs."""
    comps = components_mfcc12_0_d_a(48000.0, n_channels=2)
    p = Plan(comps, "lld", device=0)
    utts = [voiced_pcm(n, 48000, seed=300 + i, n_chan=2) for i, n in enumerate([24000, 1200, 1679, 9000])]
    pcm, off = pack_utterances(utts, n_chan=2)
    out = p.run_host(pcm, off)
    fo = p.frame_offsets(off)
    for u, x in enumerate(utts):
        ref = oracle.mfcc_d_a(x, 48000.0, n_chan=2)
        got = out[fo[u]:fo[u + 1]]
        assert got.shape == ref.shape
        assert rel_to_frame_scale(got, ref) < TOL
    p.close()


def test_batch_vs_oracle_96k_fft4096():
    """96 kHz: 2400-sample frames -> FFT 4096 (8-frame tiles, radix 16 x 16 x 8)."""
    p = Plan(components_mfcc12_0_d_a(96000.0), "lld", device=0)
    utts = [voiced_pcm(n, 96000, seed=400 + i) for i, n in enumerate([48000, 2400, 5000])]
    pcm, off = pack_utterances(utts)
    out = p.run_host(pcm, off)
    fo = p.frame_offsets(off)
    for u, x in enumerate(utts):
        ref = oracle.mfcc_d_a(x, 96000.0)
        got = out[fo[u]:fo[u + 1]]
        assert got.shape == ref.shape
        assert rel_to_frame_scale(got, ref) < TOL
    p.close()


def test_extreme_inputs(plan16):
    # silence (log floor path), full-scale square wave, single impulse
    n = 400 + 160 * 40
    sil = np.zeros(n, np.int16)
    sq = (np.where((np.arange(n) // 40) % 2 == 0, 32767, -32768)).astype(np.int16)
    imp = np.zeros(n, np.int16); imp[1234] = 32767
    pcm, off = pack_utterances([sil, sq, imp])
    out = plan16.run_host(pcm, off)
    fo = plan16.frame_offsets(off)
    assert np.isfinite(out).all()
    for u, x in enumerate([sil, sq, imp]):
        ref = oracle.mfcc_d_a(x, 16000.0)
        assert rel_to_frame_scale(out[fo[u]:fo[u + 1]], ref) < TOL


def test_device_resident_entry_point_and_determinism(plan16):
    import torch
    utts = [voiced_pcm(80240, 16000, seed=300 + i) for i in range(4)]
    pcm, off = pack_utterances(utts)
    host = plan16.run_host(pcm, off)
    d_pcm = torch.from_numpy(pcm).cuda()
    a = plan16.run_device(d_pcm, off)
    b = plan16.run_device(d_pcm, off)
    torch.cuda.synchronize()
    assert torch.equal(a, b)                                  # run-to-run bit identical
    assert np.array_equal(a.cpu().numpy(), host)              # host and device entry points agree
    assert plan16.last_launch_count() >= 1
    assert plan16.last_kernel_ms() > 0


def test_full_size_properties(plan16):
    """BASELINE cfg 2 scale (a slice of it: 200 x 500 frames), size-independent properties:
    batching invariance (an utterance's rows do not depend on its neighbours) and linearity of
    the regression stages."""
    import torch
    n_utt, L = 200, 80240
    base = voiced_pcm(L * 4, 16000, seed=999)
    pcm = np.tile(base, n_utt // 4)
    off = np.arange(n_utt + 1, dtype=np.int64) * L
    out = plan16.run_host(pcm, off)
    assert out.shape == (n_utt * 500, 39)
    blk = out.reshape(n_utt // 4, 4 * 500, 39)
    assert np.array_equal(blk[0], blk[-1]) and np.array_equal(blk[0], blk[len(blk) // 2])
    single = plan16.run_host(base[:L], np.array([0, L], np.int64))
    assert np.array_equal(single, out[:500])


def test_one_long_utterance_600s(plan16):
    """SURVEY.md 8(a') known answer: 600 s at 16 kHz -> 59 998 MFCC rows.  One utterance spanning many
    chunks of many tiles: compared with the oracle on the whole signal."""
    L = 9_600_000
    pcm = np.tile(voiced_pcm(160_000, 16000, seed=77), L // 160_000)
    off = np.array([0, L], np.int64)
    assert int(plan16.frame_offsets(off)[-1]) == 59_998
    out = plan16.run_host(pcm, off)
    ref = oracle.mfcc_d_a(pcm, 16000.0)
    assert out.shape == ref.shape == (59_998, 39)
    assert rel_to_frame_scale(out, ref) < TOL


def test_fused_and_two_kernel_paths_agree(plan16, monkeypatch):
    """The delta stages run fused inside lld_kernel by default; OSM_B200_NO_FUSE=1 selects the
    generic post_kernel path.  Both must give bit-identical rows (same float arithmetic), on
    ragged batches including 1-3 frame utterances (tick-order edge model) and long utterances
    that are split into several chunks."""
    lens = [400, 560, 720, 880, 1040, 16000, 400 + 160 * 600, 400 + 160 * 1100, 399, 5000]
    utts = [voiced_pcm(n, 16000, seed=400 + i) for i, n in enumerate(lens)]
    pcm, off = pack_utterances(utts)
    fused = plan16.run_host(pcm, off)
    monkeypatch.setenv("OSM_B200_NO_FUSE", "1")
    p2 = Plan(components_mfcc12_0_d_a(16000.0), "lld", device=0)
    two = p2.run_host(pcm, off)
    assert p2.last_launch_count() == 2 and plan16.last_launch_count() == 1
    p2.close()
    assert np.array_equal(fused, two)
    fo = plan16.frame_offsets(off)
    for u, x in enumerate(utts):
        ref = oracle.mfcc_d_a(x, 16000.0)
        if ref.shape[0]:
            assert rel_to_frame_scale(fused[fo[u]:fo[u + 1]], ref) < TOL


def test_pipelined_host_path_matches_device_path(plan16):
    """run_host cuts large batches into pieces that overlap H2D / kernels / D2H on three streams;
    rows must be bit-identical to the single-launch device path (112 MB of PCM -> 4 pieces)."""
    import torch
    n_utt, L = 700, 80240
    base = np.concatenate([voiced_pcm(L, 16000, seed=500 + i) for i in range(7)])
    pcm = np.tile(base, n_utt // 7)
    off = np.arange(n_utt + 1, dtype=np.int64) * L
    h_pcm = torch.from_numpy(pcm).pin_memory()
    h_out = torch.empty((n_utt * 500, 39), dtype=torch.float32).pin_memory()
    plan16.run_host(h_pcm, off, out=h_out)
    dev = plan16.run_device(h_pcm.cuda(), off)
    torch.cuda.synchronize()
    assert torch.equal(dev.cpu(), h_out)
    ref = oracle.mfcc_d_a(base[:L], 16000.0)
    assert rel_to_frame_scale(h_out[:500].numpy(), ref) < TOL
    assert rel_to_frame_scale(h_out[-3500:-3000].numpy(), ref) < TOL


# ---------------------------------------------------------------- PLP_0_D_A (BASELINE cfg 5)
def test_plp_goldens_and_oracle():
    from opensmile_b200 import components_plp_0_d_a
    g = np.load(os.path.join(GOLD, "plp_goldens.npz"))
    ex = np.load(os.path.join(GOLD, "mfcc_example_44k1.npz"))
    p = Plan(components_plp_0_d_a(44100.0), "lld", device=0)
    assert p.num_elements == 18
    out = p.run_host(ex["pcm"], np.array([0, ex["pcm"].size], np.int64))
    assert out.shape == (202, 18)
    assert rel_to_frame_scale(out, g["example_lld"]) < TOL
    p.close()
    # 44.1 kHz stereo (monoMixdown), ragged batch, vs the reference golden and the oracle
    p2 = Plan(components_plp_0_d_a(44100.0, n_channels=2), "lld", device=0)
    a = voiced_pcm(44100, 44100, seed=2, n_chan=2)
    others = [voiced_pcm(n, 44100, seed=600 + i, n_chan=2) for i, n in enumerate([1103, 30011, 1102, 9000])]
    pcm, off = pack_utterances([a] + others, n_chan=2)
    out2 = p2.run_host(pcm, off)
    fo = p2.frame_offsets(off)
    assert rel_to_frame_scale(out2[fo[0]:fo[1]], g["stereo44k1_lld"]) < TOL
    for u, x in enumerate([a] + others):
        ref = oracle.plp_d_a(x, 44100.0, n_chan=2)
        assert out2[fo[u]:fo[u + 1]].shape == ref.shape
        if ref.shape[0]:
            assert rel_to_frame_scale(out2[fo[u]:fo[u + 1]], ref) < TOL
    p2.close()


def test_plp_16k_batch_vs_oracle():
    from opensmile_b200 import components_plp_0_d_a
    p = Plan(components_plp_0_d_a(16000.0), "lld", device=0)
    lens = [16000, 400, 560, 80240, 5000]
    utts = [voiced_pcm(n, 16000, seed=700 + i) for i, n in enumerate(lens)]
    pcm, off = pack_utterances(utts)
    out = p.run_host(pcm, off)
    fo = p.frame_offsets(off)
    assert np.isfinite(out).all()
    for u, x in enumerate(utts):
        ref = oracle.plp_d_a(x, 16000.0)
        assert rel_to_frame_scale(out[fo[u]:fo[u + 1]], ref) < TOL
    assert p.element_names[0] == "PlpCC[0]" and p.element_names[6] == "PlpCC_de[0]"
    p.close()


# ---------------------------------------------------------------- cSpectral / cEnergy / cMZcr (general, non-fused path)
def _check_cols(got, ref, rtol=1e-5, exact_cols=()):
    """per-column check: |got-ref| <= rtol * max|ref[:,c]| ; columns in exact_cols must be equal
    (roll-off points are bin frequencies: a 1e-7 FFT perturbation must not move them here)."""
    assert got.shape == ref.shape
    for c in range(ref.shape[1]):
        scale = max(float(np.abs(ref[:, c]).max()), 1e-30)
        err = float(np.abs(got[:, c] - ref[:, c]).max())
        if c in exact_cols:
            assert np.array_equal(got[:, c], ref[:, c]), (c, err)
        else:
            assert err <= rtol * scale, (c, err, scale)


def test_spectral_compare16_and_gemaps_vs_oracle():
    from opensmile_b200 import comp, components_frontend
    utts = [voiced_pcm(n, 16000, seed=800 + i) for i, n in enumerate([16000, 320, 5000, 48000])]
    pcm, off = pack_utterances(utts)
    fe = oracle.frontend(16000.0, 0.020, 0.010, "ham")
    # ComParE is13_spectral (config/compare16/ComParE_2016_core.lld.conf.inc:283-302)
    cs = components_frontend(16000.0, 0.020, win="ham") + [
        comp("cSpectral", "spec", "mag", "spec", bands=[(250, 650), (1000, 4000)], rollOff=[0.25, 0.5, 0.75, 0.9],
             flux=1, centroid=1, maxPos=0, minPos=0, entropy=1, variance=1, skewness=1, kurtosis=1, slope=1,
             harmonicity=1, sharpness=1)]
    p = Plan(cs, "spec", device=0)
    assert p.num_elements == 15
    assert p.element_names[0] == "pcm_fftMag_fband250-650" and p.element_names[2] == "pcm_fftMag_spectralRollOff25.0"
    assert p.element_names[6] == "pcm_fftMag_spectralFlux" and p.element_names[13] == "pcm_fftMag_psySharpness"
    out = p.run_host(pcm, off)
    fo = p.frame_offsets(off)
    for u, x in enumerate(utts):
        ref = oracle.spectral(x, fe, oracle.compare16_spectral())
        _check_cols(out[fo[u]:fo[u + 1]], ref, rtol=2e-5)
    p.close()
    # GeMAPS log-spectral descriptors (config/gemaps/v01b/GeMAPSv01b_core.lld.conf.inc:321-346)
    cs = components_frontend(16000.0, 0.020, win="ham") + [
        comp("cSpectral", "lspec", "mag", "lspec", slopes=[(0, 500), (500, 1500)], flux=0, centroid=0, maxPos=0, minPos=0,
             alphaRatio=1, hammarbergIndex=1, normBandEnergies=1, squareInput=1, useLogSpectrum=1,
             freqRange=(0, 5000), oldSlopeScale=0)]
    p = Plan(cs, "lspec", device=0)
    assert p.element_names == ["pcm_fftMag_logSpectralSlopeOfBand0-500", "pcm_fftMag_logSpectralSlopeOfBand500-1500",
                               "pcm_fftMag_alphaRatioDB", "pcm_fftMag_hammarbergIndexDB"]
    out = p.run_host(pcm, off)
    # log-spectral slopes / ratios are differences of dB values of bins far below the frame peak:
    # the FFT's 2e-7-of-peak noise is a >1e-5 RELATIVE perturbation of those bins, so these four
    # columns are ill-conditioned (the float64 oracle itself sits at 1.2e-5 from the reference);
    # they are checked at 1e-4 of the column scale
    for u, x in enumerate(utts):
        ref = oracle.spectral(x, fe, oracle.gemaps_logspectral())
        _check_cols(out[fo[u]:fo[u + 1]], ref, rtol=1e-4)
    p.close()


def test_energy_zcr_bit_exact_and_multi_stream_concat():
    """cEnergy on the 20 ms framer level + cMZcr on the 60 ms framer level + windowed cEnergy, joined by
    cVectorConcat, smoothed by cContourSmoother / cDeltaRegression: time-domain ops are pure float
    arithmetic in the reference's order, so they must match the oracle bit for bit."""
    from opensmile_b200 import comp, components_frontend
    T = 16000
    utts = [voiced_pcm(n, T, seed=900 + i) for i, n in enumerate([16000, 960, 7000, 959, 31000])]
    pcm, off = pack_utterances(utts)
    cs = components_frontend(16000.0, 0.020, win="ham", prefix="a", with_fft=False)
    cs += components_frontend(16000.0, 0.060, win="gau", prefix="b", with_fft=False)[1:]
    cs += [comp("cEnergy", "e25", "aframe", "e25", rms=1, log=0),
           comp("cMZcr", "z60", "bframe", "z60", zcr=1, mcr=1, amax=1, maxmin=1, dc=1),
           comp("cEnergy", "e60", "bwin", "e60", rms=1, log=1, energy2=1),
           comp("cContourSmoother", "sm", "e25", "e25s", smaWin=3),
           comp("cDeltaRegression", "de", "e25s", "e25sd", deltawin=2),
           comp("cVectorConcat", "cat", "e25;z60;e60;e25s;e25sd", "lld", includeSingleElementFields=1)]
    p = Plan(cs, "lld", device=0)
    assert p.element_names == ["pcm_RMSenergy", "pcm_zcr", "pcm_mcr", "pcm_absmax", "pcm_max", "pcm_min", "pcm_dc",
                               "pcm_RMSenergy", "pcm_SQUAREDenergy", "pcm_LOGenergy", "pcm_RMSenergy_sma", "pcm_RMSenergy_sma_de"]
    out = p.run_host(pcm, off)
    fo = p.frame_offsets(off)
    fa = oracle.frontend(16000.0, 0.020, 0.010, "ham")
    fb = oracle.frontend(16000.0, 0.060, 0.010, "gau", sigma=0.4)
    for u, x in enumerate(utts):
        n60 = oracle.geometry(fb, len(x))[3]
        assert fo[u + 1] - fo[u] == max(n60, 0)           # concat = min over inputs (60 ms stream is shortest)
        if n60 <= 0:
            continue
        e25 = oracle.energy(x, fa, oracle.Energy(0, 1, 0, 0, 1, 1, 1, 0, 0, 0), 0)
        z60 = oracle.mzcr(x, fb, oracle.MZcr(1, 1, 1, 1, 1), 0)
        e60 = oracle.energy(x, fb, oracle.Energy(0, 1, 1, 1, 1, 1, 1, 0, 0, 0), 1)
        sm, c0 = oracle.sma_chained(e25, 3, e25.shape[0])
        de, _ = oracle.delta_chained(sm, 2, c0)
        ref = np.concatenate([e25[:n60], z60[:n60], e60[:n60], sm[:n60], de[:n60]], axis=1)
        got = out[fo[u]:fo[u + 1]]
        assert np.array_equal(got, ref), (u, np.abs(got - ref).max(axis=0))
    p.close()


def test_acf_pitchacf_vs_oracle():
    """cAcf (ACF) + cAcf (cepstrum) -> cPitchACF (config/prosody/prosodyAcf.conf shape): per-frame
    voicing / HNR and the per-utterance F0 smoothing state machine.  F0 candidates are lags
    (integers): the F0 / F0raw / F0env columns must match the oracle exactly unless a peak decision
    flips, which is counted and bounded."""
    from opensmile_b200 import comp, components_frontend
    for fs, sr in ((0.050, 16000), (0.025, 16000)):
        utts = [voiced_pcm(n, sr, seed=950 + i) for i, n in enumerate([32000, 900, 12000, 799])]
        pcm, off = pack_utterances(utts)
        cs = components_frontend(float(sr), fs, win="gau", sigma=0.4, zero_pad_symmetric=0) + [
            comp("cAcf", "acf", "mag", "acf"),
            comp("cAcf", "cep", "mag", "cepstrum", cepstrum=1, usePower=0),
            comp("cPitchACF", "pitch", "acf;cepstrum", "pitch", maxPitch=500.0, voiceProb=1, voiceQual=1, HNR=1, HNRdB=1,
                 linHNR=1, F0=1, F0raw=1, F0env=1, voicingCutoff=0.55)]
        p = Plan(cs, "pitch", device=0)
        assert p.element_names == ["voiceProb", "HNR", "HNRdBacf", "linearHNRacf", "voiceQual", "F0", "F0raw", "F0env"]
        out = p.run_host(pcm, off)
        fo = p.frame_offsets(off)
        fe = oracle.frontend(float(sr), fs, 0.010, "gau", sigma=0.4, zero_pad_symmetric=0)
        cfg = oracle.pitchacf_cfg(voiceProb=1, voiceQual=1, HNR=1, HNRdB=1, linHNR=1, F0=1, F0raw=1, F0env=1)
        for u, x in enumerate(utts):
            ref = oracle.pitchacf(x, fe, cfg)
            got = out[fo[u]:fo[u + 1]]
            assert got.shape == ref.shape
            if not ref.shape[0]:
                continue
            flips = int((got[:, 6] != ref[:, 6]).sum())            # F0raw = 1/(lag*Ts): differs only if the peak lag flips
            assert flips <= max(1, ref.shape[0] // 100), flips
            if flips == 0:
                assert np.array_equal(got[:, 5:8], ref[:, 5:8])
                for c in range(5):
                    scale = max(float(np.abs(ref[:, c]).max()), 1e-6)
                    assert float(np.abs(got[:, c] - ref[:, c]).max()) <= 1e-4 * scale, c
        p.close()
