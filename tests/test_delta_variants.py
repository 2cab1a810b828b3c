"""SURVEY.md 8(a) row a-13, the variants of cDeltaRegression (dspcore/deltaRegression.cpp:100-168): relativeDelta, absOutput,
halfWaveRect (which wins over absOutput), another window.  Goldens: the unmodified reference on tests/configs/delta_variants.conf
([MFCC 1..12 | their regression], truncated to the statics' length by the concat).  CPU: the oracle's stage on the reference's own
statics is bit-identical to the reference's regression columns.  GPU: the configuration through the session."""
import os

import numpy as np
import pytest

from oracle import oracle
from opensmile_b200.synth import mixed_pcm

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "delta_variants.npz"))
CONF = os.path.join(HERE, "configs", "delta_variants.conf")
# name -> (session options, (window, relative, absOutput, halfWaveRect))
VARIANTS = {"plain": ({}, (2, 0, 0, 0)), "relative": ({"relativeDelta": "1"}, (2, 1, 0, 0)), "abs": ({"absOutput": "1"}, (2, 0, 1, 0)),
            "halfwave": ({"halfWaveRect": "1", "absOutput": "1"}, (2, 0, 1, 1)),
            "rel_abs_w3": ({"relativeDelta": "1", "absOutput": "1", "deltawin": "3"}, (3, 1, 1, 0))}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_stage_on_the_reference_statics_is_bit_identical(name):
    W, rel, ab, hw = VARIANTS[name][1]
    ref = G["rows_" + name]
    stat, de = ref[:, :12], ref[:, 12:]
    got = oracle.delta_variant(stat, W, rel, ab, hw)[:len(stat)]
    assert np.array_equal(got.view(np.uint32), de.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_session_equals_the_reference(name):
    from opensmile_b200.session import Session
    pcm = mixed_pcm(16000, 16000, seed=4)
    s = Session(CONF, options=dict(VARIANTS[name][0], O="x.htk"), device=0)
    rows, fo = s.extract_pcm(pcm, np.array([0, pcm.size], np.int64), 16000.0, 1)
    s.close()
    ref = G["rows_" + name]
    assert rows.shape == ref.shape
    # the statics to the MFCC rule; the regression columns on what the kernel computed from ITS statics: the stage itself is exact
    W, rel, ab, hw = VARIANTS[name][1]
    own = oracle.delta_variant(rows[:, :12], W, rel, ab, hw)[:len(rows)]
    assert np.array_equal(rows[:, 12:].view(np.uint32), own.view(np.uint32))
    scale = np.abs(ref[:, :12]).max(axis=0, keepdims=True)
    assert (np.abs(rows[:, :12] - ref[:, :12]) / scale).max() < 5e-5
    if not rel:                                     # a relative difference divides by statics that can be ~0: compared through `own` above
        dscale = np.abs(ref[:, 12:]).max(axis=0, keepdims=True)
        assert (np.abs(rows[:, 12:] - ref[:, 12:]) / dscale).max() < 5e-5
