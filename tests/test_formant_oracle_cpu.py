"""Formant / harmonics chain of the GeMAPS graphs: the numpy restatement (oracle/formant_oracle.py)
against level taps and LLD files of the UNMODIFIED reference (tests/golden/formant_goldens.npz,
scripts/make_golden_formant.py).  This pins the checker; the product's kernels (formant.cu, harmonics.cu) are held to the same
taps through a host build of their statements in tests/test_formant_kernel_cpu.py."""
import os

import numpy as np

from opensmile_b200.synth import mixed_pcm
from oracle import formant_oracle as fo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "formant_goldens.npz"))


def test_resampled_frames():
    fmt, res, lpcs = fo.gemaps_formant_chain(mixed_pcm(24000, 16000, seed=3), taps=True)
    assert res.shape == G["res"].shape
    scale = np.abs(G["res"]).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(res - G["res"]) / scale).max() < 2e-6          # own FFT vs Ooura's float FFT


def test_lpc_and_formants_exact_given_inputs():
    """float autocorrelation + Durbin and the root -> formant mapping reproduce the reference bit for bit"""
    a = np.stack([fo.lpc_acf(fo.autocorr(x, 12), 11)[0] for x in G["res"]])
    assert np.array_equal(a, G["lpc"])
    f = np.stack([np.concatenate(fo.formants_from_lpc(x, 1.0 / 11000.0, 5, 50.0, 5450.0)) for x in G["lpc"]])
    assert (np.abs(f - G["fmt"]) / (np.abs(G["fmt"]) + 1.0)).max() < 1e-6


def test_conditioning_of_the_chain_is_documented():
    """Order-11 LPC in float amplifies a 2e-7 perturbation of the frame (own FFT vs Ooura) to percent-level
    changes of the coefficients on strongly harmonic frames: end-to-end parity of the formant columns needs a
    bit-identical FFT + resampler in front of cLpc."""
    fmt, res, lpcs = fo.gemaps_formant_chain(mixed_pcm(24000, 16000, seed=3), taps=True)
    err = np.abs(lpcs - G["lpc"]).max(axis=1)
    assert np.median(err) < 1e-2 and err.max() > 1e-3


def test_harmonics_exact_given_inputs():
    """cHarmonics (HNR from the ACF, H1-H2, H1-A3, formant amplitudes) on the reference's own F0 / formant / magnitude
    levels: every decision (harmonic peak search, formant-range maximum) and value reproduced"""
    T = G["h_harm"].shape[0]
    frq = np.arange(513, dtype=np.float64) * (1.0 / 0.064)          # bin k -> k / frameSizeSec (dspcore/transformFft.cpp:111-115)
    got = np.stack([fo.harmonics_gemaps(G["h_f0"][t, 0], G["h_fmt"][t, :5], G["h_mag"][t], frq) for t in range(T)])
    assert np.abs(got - G["h_harm"][:T]).max() < 1e-4
    voiced = G["h_f0"][:T, 0] > 0
    assert voiced.sum() > 20 and (~voiced).sum() > 20


def test_whole_chain_bit_identical_with_the_reference_fft():
    """With the reference's own FFT (oracle/_ref/libfftsg.so = src/dspcore/fftsg.c compiled where it lies) in front, the
    restated chain PCM -> window -> FFT -> cSpecResample -> cLpc -> cFormantLpc is bit-identical to the reference end
    to end: everything behind the FFT is exact, so the FFT's butterfly order is what the product has to reproduce."""
    import pytest
    if not fo.ref_fft_available():
        pytest.skip("oracle/_ref/libfftsg.so not built (make -C oracle ref)")
    fmt, res, lpcs = fo.gemaps_formant_chain(mixed_pcm(24000, 16000, seed=3), taps=True, exact_fft=True)
    assert np.array_equal(res, G["res"])
    assert np.array_equal(lpcs, G["lpc"])
    assert (np.abs(fmt - G["fmt"]) / (np.abs(G["fmt"]) + 1.0)).max() < 1e-6


def test_gemaps_voice_quality_levels_end_to_end():
    """PCM -> the four voice-quality levels of the shipped GeMAPS graph (pitch with semitone scale, jitter / shimmer dB,
    formants, harmonics): pitch and jitter match with the oracle's own FFT; formants and the harmonic columns that read
    them need the bit-identical FFT in front of cLpc (then everything matches)"""
    import pytest
    pcm = mixed_pcm(24000, 16000, seed=3)

    def rel(a, b):
        return float((np.abs(a - b) / (np.abs(b).max(axis=0) + 1e-30)).max())

    pitch, jit, fmt, harm = fo.gemaps_vq_levels(pcm, exact_fft=False)
    assert rel(pitch, G["g_f0"]) < 2e-6 and np.array_equal(jit, G["g_jit"])
    assert rel(harm[:, :2], G["g_harm"][:, :2]) < 1e-5            # HNR from the ACF and H1-H2 do not read the formants
    if not fo.ref_fft_available():
        pytest.skip("oracle/_ref/libfftsg.so not built (make -C oracle ref)")
    pitch, jit, fmt, harm = fo.gemaps_vq_levels(pcm, exact_fft=True)
    assert rel(fmt, G["g_fmt"]) < 1e-6 and rel(harm, G["g_harm"]) < 1e-5 and rel(pitch, G["g_f0"]) < 2e-6


def test_shipped_gemaps_lld_level_end_to_end():
    """The complete `lld` level of the shipped config/gemaps/v01b/GeMAPSv01b.conf (18 columns incl. the selectors' element
    order, sma3 / sma3nz smoothing and the rows smoothed while the jitter level lags at the end of input), PCM to rows,
    against the reference's -lldhtkoutput file"""
    import pytest
    if not fo.ref_fft_available():
        pytest.skip("oracle/_ref/libfftsg.so not built (make -C oracle ref)")
    for key, pcm in (("gemaps_lld_m24k", mixed_pcm(24000, 16000, seed=3)), ("gemaps_lld_m40k", mixed_pcm(40000, 16000, seed=5))):
        got, ref = fo.gemaps_lld(pcm, exact_fft=True), G[key]
        assert got.shape == ref.shape == (ref.shape[0], 18)
        assert (np.abs(got - ref) / (np.abs(ref).max(axis=0) + 1e-30)).max() < 5e-6


def test_shipped_egemaps_lld_level_end_to_end():
    """The `lld` level of the shipped config/egemaps/v02/eGeMAPSv02.conf (25 columns: GeMAPS plus spectral flux from the
    flux-only cSpectral instance, MFCC 1-4 and the F2 / F3 bandwidths; every voice-quality column lags behind the selector in
    front of the smoother) against the reference's -lldhtkoutput file, column names against its CSV header"""
    import pytest
    if not fo.ref_fft_available():
        pytest.skip("oracle/_ref/libfftsg.so not built (make -C oracle ref)")
    assert list(G["names_egemaps_lld"]) == fo.EGEMAPS_LLD_NAMES
    for key, pcm in (("egemaps_lld_m24k", mixed_pcm(24000, 16000, seed=3)), ("egemaps_lld_m40k", mixed_pcm(40000, 16000, seed=5))):
        got, ref = fo.egemaps_lld(pcm, exact_fft=True), G[key]
        assert got.shape == ref.shape == (ref.shape[0], 25)
        assert (np.abs(got - ref) / (np.abs(ref).max(axis=0) + 1e-30)).max() < 5e-6


def test_flux_only_spectral_instance():
    """[egemapsv02_spectral_flux] requests nothing but the magnitude spectrum: one column, first frame 0"""
    from oracle import oracle
    pcm = mixed_pcm(8000, 16000, seed=1)
    fe = oracle.frontend(16000, 0.020, 0.010, win="ham")
    x = oracle.spectral(pcm, fe, oracle.spectral_cfg(flux=1, centroid=0, maxPos=0, minPos=0, normBandEnergies=1, squareInput=1,
                                                      useLogSpectrum=1, freqRangeLo=0, freqRangeHi=5000, oldSlopeScale=0))
    assert x.shape[1] == 1 and x[0, 0] == 0.0 and np.all(x[1:, 0] > 0)


def test_end_to_end_bound_of_the_formant_dependent_columns():
    """What an implementation with a different (equally accurate) FFT / resampler in front of cLpc can reach on the shipped
    eGeMAPSv02 LLD level: columns that do not read the formant chain agree to 2e-6, the formant-dependent ones (F1-F3
    frequency / bandwidth / amplitude, H1-A3) to 1e-5 in the median with ~10 % of the rows beyond 1e-3 (order-11 float LPC,
    DESIGN.md 3.6).  tests/test_zzz_gemaps_gpu.py holds the GPU rows to this bound with a factor-2 margin."""
    names = list(G["names_egemaps_lld"])
    fdep = [i for i, n in enumerate(names) if n.startswith(("F1", "F2", "F3")) or "H1-A3" in n]
    rest = [i for i in range(len(names)) if i not in fdep]
    pcm = mixed_pcm(24000, 16000, seed=3)
    got, ref = fo.egemaps_lld(pcm, exact_fft=False), G["egemaps_lld_m24k"]
    err = np.abs(got - ref) / (np.abs(ref).max(axis=0) + 1e-30)
    assert err[:, rest].max() < 5e-6
    assert np.median(err[:, fdep]) < 5e-5 and (err[:, fdep].max(axis=1) > 1e-3).mean() < 0.15


def test_shipped_gemaps_v01a_lld_level():
    """config/gemaps/v01a/GeMAPSv01a.conf differs from v01b in zeroPadSymmetric = 0 on both FFTs (the formant branch sees
    another phase), maxF = 5500 (= Nyquist of the 11 kHz frames) and useBrokenJitterThresh = 1: same oracle, three
    switches, against the reference's rows (tests/golden/gemaps_family.npz)"""
    import pytest
    if not fo.ref_fft_available():
        pytest.skip("oracle/_ref/libfftsg.so not built (make -C oracle ref)")
    R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gemaps_family.npz"))
    pcm = mixed_pcm(24000, 16000, seed=3)
    for key, v in (("GeMAPSv01a", True), ("GeMAPSv01b", False)):
        got, ref = fo.gemaps_lld(pcm, exact_fft=True, v01a=v), R[key]
        assert got.shape == ref.shape
        assert (np.abs(got - ref) / (np.abs(ref).max(axis=0) + 1e-30)).max() < 5e-6
    assert not np.array_equal(R["GeMAPSv01a"], R["GeMAPSv01b"])


def test_shipped_egemaps_v01_lld_levels():
    """eGeMAPSv01a.conf / eGeMAPSv01b.conf (23 columns) -- with GeMAPSv01a/b and eGeMAPSv02 the oracle now reproduces the LLD
    level of all five shipped feature-set files of the family"""
    import pytest
    if not fo.ref_fft_available():
        pytest.skip("oracle/_ref/libfftsg.so not built (make -C oracle ref)")
    R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gemaps_family.npz"))
    pcm = mixed_pcm(24000, 16000, seed=3)
    for key, v in (("eGeMAPSv01a", True), ("eGeMAPSv01b", False)):
        got, ref = fo.egemaps_v01_lld(pcm, exact_fft=True, v01a=v), R[key]
        assert got.shape == ref.shape == (ref.shape[0], 23)
        assert (np.abs(got - ref) / (np.abs(ref).max(axis=0) + 1e-30)).max() < 5e-6
