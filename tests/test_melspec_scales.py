"""SURVEY.md 8(a) row a-7, the frequency scales besides mel (lldcore/melspec.cpp:100-135; smileutil/smileUtil.c:1097-1204): bark,
bark_speex, bark_schroed, semitone, linear, log.  Goldens: the unmodified reference's MFCC 0..12 behind cMelspec on each scale
(tests/configs/mfcc_scales.conf, scripts/make_golden_melspec_scales.py).  CPU: the oracle's filter design on each scale against those
rows.  GPU: the same configuration file through the session and the C ABI."""
import os

import numpy as np
import pytest

from oracle import oracle
from opensmile_b200.synth import mixed_pcm

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "melspec_scales.npz"))
CONF = os.path.join(HERE, "configs", "mfcc_scales.conf")
# name -> (session options, oracle spec_scale, scale_param, lofreq)
VARIANTS = {"mel": ({"scale": "mel"}, 0, 0.0, 20.0), "bark": ({"scale": "bark"}, 1, 0.0, 20.0), "bark_speex": ({"scale": "bark_speex"}, 2, 0.0, 20.0),
            "bark_schroed": ({"scale": "bark_schroed"}, 3, 0.0, 20.0),
            "semitone": ({"scale": "semitone", "firstNote": "55.0", "lofreq": "60"}, 4, 55.0, 60.0), "linear": ({"scale": "linear"}, 5, 0.0, 20.0),
            "log2": ({"scale": "log", "lofreq": "50"}, 6, 2.0, 50.0), "log10": ({"scale": "log", "logScaleBase": "10.0", "lofreq": "50"}, 6, 10.0, 50.0)}


def _close(got, ref):
    from conftest import column_scale_report
    worst, share = column_scale_report(got, ref)
    assert worst < 5e-5 and share <= 2e-3, (worst, share)


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_filter_design_on_every_scale(name):
    _, scale, param, lofreq = VARIANTS[name]
    fe = oracle.Frontend(16000.0, 0.025, 0.010, 0, 0.0, oracle.WIN["ham"], 0.4, 1.0, 0.0, 0)
    ms = oracle.Melspec(26, lofreq, 8000.0, 1, 0, scale, param)
    mf = oracle.Mfcc(0, 12, 22.0, 1e-8, 0)
    got = oracle.mfcc_d_a(mixed_pcm(24000, 16000, seed=3), 16000, cfg=(fe, ms, mf))[:, :13]
    ref = G["mfcc_" + name]
    assert got.shape == ref.shape
    _close(got, ref)


def test_scale_names_parse_like_the_reference():
    """case-insensitive, semi / lin / log by prefix, unknown names fall back to mel (melspec.cpp:100-126); htkcompatible = 1 forces mel"""
    from opensmile_b200.session import Session
    for opt, want in (("Bark", 1), ("bark_speex", 2), ("BARK_SCHROED", 3), ("semitones", 4), ("linear", 5), ("lin", 5), ("logarithmic", 6), ("nonsense", 0)):
        s = Session(CONF, options={"scale": opt}, device=-1)
        comps, _ = s.components(16000.0, 1)
        ms = [c for c in comps if c.name == b"melspec"][0]
        assert ms.u.melspec.specScale == want, (opt, ms.u.melspec.specScale)
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_session_on_every_scale_equals_the_reference(name):
    from opensmile_b200.session import Session
    opts = VARIANTS[name][0]
    pcm = mixed_pcm(24000, 16000, seed=3)
    s = Session(CONF, options=dict(opts), device=0)
    rows, fo = s.extract_pcm(pcm, np.array([0, pcm.size], np.int64), 16000.0, 1)
    s.close()
    ref = G["mfcc_" + name]
    assert rows.shape == ref.shape
    _close(rows, ref)


def test_band_level_as_output_level_names():
    """cMelspec as the level a sink reads (log-mel spectrogram style graphs): nBands elements named like the magnitude field
    (lldcore/melspec.cpp:175-178: no default nameAppend)"""
    from opensmile_b200.session import Session
    s = Session(CONF, output_level="melspec", device=-1)
    assert s.element_names() == [str(x) for x in G["melspec_names"]]
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,opts", [("mel", {"scale": "mel"}), ("bark", {"scale": "bark"}), ("htk", {"melhtk": "1"})])
def test_band_level_as_output_level(name, opts):
    """the band sums themselves (two-tap filterbank, double product per bin, htk scaling 32767^2) against the reference's melspec
    level: every band within 1e-5 of its scale"""
    from opensmile_b200.session import Session
    pcm = mixed_pcm(24000, 16000, seed=3)
    s = Session(CONF, options=dict(opts), output_level="melspec", device=0)
    rows, fo = s.extract_pcm(pcm, np.array([0, pcm.size], np.int64), 16000.0, 1)
    s.close()
    ref = G["melspec_" + name]
    assert rows.shape == ref.shape
    err = np.abs(rows - ref) / np.abs(ref).max(axis=0, keepdims=True)
    assert err.max() < 1e-5, float(err.max())


@pytest.mark.parametrize("name,scale,htk", [("mel", 0, 0), ("bark", 1, 0), ("htk", 0, 1)])
def test_oracle_band_level_equals_the_reference_melspec_level(name, scale, htk):
    """the oracle's band sums (tap of oracle.mfcc_d_a) against the reference's melspec level, incl. the htk scaling 32767^2"""
    fe = oracle.Frontend(16000.0, 0.025, 0.010, 0, 0.0, oracle.WIN["ham"], 0.4, 1.0, 0.0, 0)
    ms = oracle.Melspec(26, 20.0, 8000.0, 1, htk, scale, 0.0)
    mf = oracle.Mfcc(0, 12, 22.0, 1e-8, 0)
    _, _, mel = oracle.mfcc_d_a(mixed_pcm(24000, 16000, seed=3), 16000, cfg=(fe, ms, mf), taps=True)
    ref = G["melspec_" + name]
    assert mel.shape == ref.shape
    err = np.abs(mel - ref) / np.abs(ref).max(axis=0, keepdims=True)
    assert err.max() < 1e-5, float(err.max())
