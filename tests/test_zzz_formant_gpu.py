"""Formant chain on the GPU (cWindower -> cTransformFFT -> cSpecResample -> cLpc -> cFormantLpc as one kernel,
opensmile_b200/csrc/formant.cu) through the session C ABI.

At FFT size 512 (16 kHz, 20 ms frames: every GeMAPS-family configuration) the kernel transforms with the reference's rounding
sequence (fft_ref_order.cuh) and resamples with the reference's float inverse sum, so the resampled frames are bit-identical to
the reference's and the formants follow: the bar against the reference's own formant level is 1e-5 of each column's scale on
EVERY row (the roots differ from the reference's QR iteration at the 1e-7 level only).  The kernel is also held to 1e-6 against
the host build of the same statements (tests/formant_harness.py)."""
import os

import numpy as np
import pytest

import formant_harness as fh
from opensmile_b200.synth import mixed_pcm, voiced_pcm

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CONF = os.path.join(HERE, "configs", "formant_chain.conf")


def _run(pcms):
    from opensmile_b200.session import Session
    s = Session(CONF, device=0)
    off = np.concatenate([[0], np.cumsum([len(x) for x in pcms])]).astype(np.int64)
    rows, fo = s.extract_pcm(np.concatenate(pcms + [np.zeros(8, np.int16)]), off, 16000.0, 1)
    s.close()
    return [rows[fo[i]:fo[i + 1]] for i in range(len(pcms))]


def test_formant_rows_equal_the_host_build():
    pcms = [mixed_pcm(24000, 16000, seed=3), voiced_pcm(16000, 16000, seed=2), mixed_pcm(5000, 16000, seed=9),
            np.zeros(4000, np.int16), mixed_pcm(319, 16000, seed=1)]
    rows = _run(pcms)
    assert len(rows) == len(pcms)
    for pcm, got in zip(pcms, rows):
        ref = fh.formant_chain(pcm)
        assert got.shape == ref.shape
        if ref.size:
            scale = np.abs(ref).max(axis=0) + 1.0
            assert (np.abs(got - ref) / scale).max() < 1e-6


def test_formant_rows_against_the_reference_taps():
    """end to end against the reference's own cFormantLpc level: every row, every column within 1e-5"""
    G = np.load(os.path.join(HERE, "golden", "formant_goldens.npz"))
    got = _run([mixed_pcm(24000, 16000, seed=3)])[0]
    ref = G["fmt"]
    assert got.shape == ref.shape
    err = np.abs(got - ref) / np.abs(ref).max(axis=0)
    assert err.max() < 1e-5
