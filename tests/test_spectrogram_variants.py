"""SURVEY.md 8(a) row a-6, the output variants of cFFTmagphase (dspcore/fftmagphase.cpp:215-255): magnitude, normalise (spectral
density), power, normalise + power, dBpsd with its floor -- served where the level is the output level (the -dB switch of the
reference's config/spectrum/spectrogram.conf).  Goldens: the unmodified reference (scripts/make_golden_spectrogram_variants.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "spectrogram_variants.npz"))
CONF = os.path.join(HERE, "configs", "spectrogram_variants.conf")
VARIANTS = {"mag": {}, "specdens": {"normalise": "1"}, "powspec": {"power": "1"}, "powspecdens": {"normalise": "1", "power": "1"},
            "dbpsd": {"dB": "1"}, "dbpsd_floor": {"dB": "1", "dBpnorm": "60.0", "mindBp": "-20.0"}}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_names_follow_the_variant(name):
    from opensmile_b200.session import Session
    s = Session(CONF, options=dict(VARIANTS[name], O="x.htk"), device=-1)
    n = s.element_names()
    assert len(n) == 257 and n[0] == str(G["name0_" + name]) and n[-1] == str(G["name0_" + name]).replace("[0]", "[256]")
    s.close()


def test_variants_are_refused_in_front_of_a_consumer(tmp_path):
    """cMelspec & co. read the plain magnitude: a power / dB level below them is not the same graph"""
    from opensmile_b200.session import Session, SessionError
    from opensmile_b200 import capi
    txt = open(os.path.join(HERE, "configs", "mfcc_scales.conf")).read().replace("[fftmag:cFFTmagphase]", "[fftmag:cFFTmagphase]\npower = 1")
    p = tmp_path / "c.conf"
    p.write_text(txt)
    with pytest.raises(SessionError) as e:
        Session(str(p), options={"O": "x.htk"}, device=-1)
    assert e.value.status == capi.ERR_UNSUPPORTED and "output level" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_values_equal_the_reference(name):
    from opensmile_b200.session import Session
    pcm = np.ascontiguousarray(G["pcm"])
    s = Session(CONF, options=dict(VARIANTS[name], O="x.htk"), device=0)
    rows, fo = s.extract_pcm(pcm, np.array([0, pcm.size], np.int64), 16000.0, 1)
    s.close()
    ref = G["rows_" + name]
    assert rows.shape == ref.shape
    if name.startswith("dbpsd"):
        # decibels: the 2e-7-of-the-frame-peak difference between the two FFTs is a larger RELATIVE difference on a weak bin, i.e. a
        # larger dB difference there (60 dB below the peak: ~1e-3 dB); the floor is exact
        assert np.abs(rows - ref).max() < 0.02 and np.abs(rows - ref).mean() < 1e-4, (float(np.abs(rows - ref).max()), float(np.abs(rows - ref).mean()))
        floor = ref.min()
        if (ref == floor).sum() > 1:
            assert np.array_equal(rows == floor, ref == floor)
    else:
        peak = np.abs(ref).max(axis=1, keepdims=True)                  # per frame: the FFT's error scales with the frame's peak
        lin = 2.0 if "pow" in name else 1.0
        assert (np.abs(rows - ref) / peak).max() < 2e-6 * lin, float((np.abs(rows - ref) / peak).max())


def test_oracle_restatement_of_the_variants_on_the_reference_magnitudes():
    """oracle-side restatement (numpy, float32) of dspcore/fftmagphase.cpp:223-255 applied to the reference's own magnitude rows
    reproduces its variant rows (the reference squares re / im directly, so the power forms differ from |X|^2 by an ulp or two)"""
    mag = G["rows_mag"].astype(np.float32)
    N = np.float32(512.0)
    edge = np.zeros(mag.shape[1], bool); edge[0] = edge[-1] = True
    dens = (np.float32(1.0) / N) * mag
    powd = np.where(edge, dens * dens, (np.float32(1.0) / (N * N)) * (mag * mag))
    for name, got in (("specdens", dens), ("powspec", mag * mag), ("powspecdens", powd)):
        ref = G["rows_" + name]
        assert np.all(np.abs(got - ref) <= 4e-7 * np.abs(ref) + 1e-30), name
    for name, norm, floor in (("dbpsd", 90.302, -102.0), ("dbpsd_floor", 60.0, -20.0)):
        floor = max(floor, norm - 120.0)
        with np.errstate(divide="ignore"):
            v = np.where(edge, np.float32(norm) + np.float32(20.0) * np.log10(dens), np.float32(norm) + np.float32(10.0) * np.log10(powd))
        got = np.maximum(np.float32(floor), v.astype(np.float32))
        assert np.abs(got - G["rows_" + name]).max() < 2e-4, name
