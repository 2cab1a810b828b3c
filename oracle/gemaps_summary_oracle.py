"""TEST INFRASTRUCTURE ONLY (tests/, scripts/): CPU restatement of the shipped GeMAPS / eGeMAPS summary graphs.  The product path
(opensmile_b200/host/front.cpp + csrc/functionals.cu + csrc/pitch.cu) never imports this file.

What the reference does (file:line under /root/reference):
  config/gemaps/v01b/GeMAPSv01b_core.lld.conf.inc:385-433   cValbasedSelector gates (voiced / unvoiced frames) in front of the smoothers
  src/other/valbasedSelector.cpp:195-233                    the gate: idx element compared with the threshold, zeroVec -> outputVal
  config/gemaps/v01b/GeMAPSv01b_core.func.conf.inc:17-139   temporal summaries (Segments nonX / eqX, Peaks2 numPeaks) + cDataSelector renaming
  config/gemaps/v01b/GeMAPSv01b_core.func.conf.inc:141-289  F0 / loudness: Moments, Percentiles, Peaks2 slopes; voiced / unvoiced means
  config/egemaps/v02/eGeMAPSv02_core.func.conf.inc:9-100    equivalent sound level (Means.amean -> cVectorOperation dBp), MVR / MeanUV / MVRVoiced
  config/egemaps/v02/eGeMAPSv02.conf:31-34                  order of the summary row (cVectorConcat funcconcat)
  src/other/vectorOperation.cpp:508-517                     dBp = 10 / ln 10 * ln(max(x, logfloor)), float

Parity: pinned -- tests/test_functionals_cpu.py::test_gemaps_summary_oracle_on_the_reference_levels feeds the reference's own dumps of
the seven input levels (tests/golden/gemaps_func_levels.npz, written by the unmodified reference through extra cCsvSink instances,
scripts/make_golden_gemaps_func_levels.py) and compares with the reference's -csvoutput row (tests/golden/gemaps_func.npz)."""
import numpy as np

from . import functionals_oracle as fo

F32 = np.float32
SEG, SEC = fo.SEGMENT, fo.SECOND


def valbased_gate(sel, data, threshold=1e-6, invert=False, allow_equal=False, output_val=0.0):
    """other/valbasedSelector.cpp:195-233 with zeroVec = 1, removeIdx = 1: rows of `data` [T, K] whose selector value passes are
    copied, the others are set to output_val"""
    sel = np.asarray(sel, F32)
    thr = F32(threshold)
    ok = (sel < thr) if invert else (sel > thr)
    if allow_equal:
        ok = ok | (sel == thr)
    out = np.array(data, F32, copy=True)
    out[~ok] = F32(output_val)
    return out


def dbp(x, logfloor=1e-12):
    """other/vectorOperation.cpp:508-517 in float"""
    x = np.asarray(x, F32)
    factor = F32(10.0 / np.log(10.0))
    fl = F32(logfloor)
    return (factor * np.log(np.where(x > fl, x, fl).astype(F32))).astype(F32)


_P2_OFF = {k: 0 for k in fo.PEAKS2_NAMES}
_MVR = dict(variance=0, stddev=0, skewness=0, kurtosis=0, amean=1, stddevNorm=2, doRatioLimit=0)
_MEAN = dict(variance=0, stddev=0, skewness=0, kurtosis=0, amean=1, stddevNorm=0, doRatioLimit=0)
_PCTL = dict(percentile=[0.20, 0.50, 0.80], pctlrange=[(0, 2)], interp=1)
_SLOPES = dict(_P2_OFF, meanRisingSlope=1, stddevRisingSlope=1, meanFallingSlope=1, stddevFallingSlope=1, norm=SEC, norm_set=True,
               relThresh=0.1, dynRelThresh=0, doRatioLimit=0)

# GeMAPSv01b_core.func.conf.inc:141-215 (F0: non-zero values only) / :216-289 (loudness)
F0_SPEC = fo.Spec(["Moments", "Percentiles", "Peaks2"], non_zero=1, master_norm=SEG, moments=_MVR, percentiles=_PCTL, peaks2=_SLOPES)
LOUD_SPEC = fo.Spec(["Moments", "Percentiles", "Peaks2"], non_zero=0, master_norm=SEG, moments=_MVR, percentiles=_PCTL, peaks2=_SLOPES)
# eGeMAPSv02_core.func.conf.inc:51-100
MVR_SPEC = fo.Spec(["Moments"], non_zero=0, master_norm=SEG, moments=_MVR)
MVR_VOICED_SPEC = fo.Spec(["Moments"], non_zero=1, master_norm=SEG, moments=_MVR)
MEAN_UV_SPEC = fo.Spec(["Moments"], non_zero=1, master_norm=SEG, moments=_MEAN)
# GeMAPSv01b_core.func.conf.inc:46-131
_SEGS = dict(maxNumSeg=1000, X=0.0, meanSegLen=1, maxSegLen=0, minSegLen=0, segLenStddev=1, norm=SEC, norm_set=True)
VOICED_SEG_SPEC = fo.Spec(["Segments"], non_zero=0, master_norm=SEC, segments=dict(_SEGS, segmentationAlgorithm="nonX", numSegments=1))
PAUSE_SEG_SPEC = fo.Spec(["Segments"], non_zero=0, master_norm=SEC, name_append="f0pause", segments=dict(_SEGS, segmentationAlgorithm="eqX", numSegments=0))
PEAKS_SPEC = fo.Spec(["Peaks2"], non_zero=0, master_norm=SEC, peaks2=dict(_P2_OFF, numPeaks=1, norm=SEC, norm_set=True, relThresh=0.1, dynRelThresh=0))
# eGeMAPSv02_core.func.conf.inc:12-33
LEQ_SPEC = fo.Spec(["Means"], non_zero=0, master_norm=None,
                   means=dict(amean=1, absmean=0, qmean=0, nzamean=0, nzabsmean=0, nzqmean=0, nzgmean=0, nnz=0))

TEMPORAL_PICK = [("loudness_sma3_numPeaks", "loudnessPeaksPerSec"), ("F0semitoneFrom27.5Hz_sma3nz_numSegments", "VoicedSegmentsPerSec"),
                 ("F0semitoneFrom27.5Hz_sma3nz_meanSegLen", "MeanVoicedSegmentLengthSec"),
                 ("F0semitoneFrom27.5Hz_sma3nz_segLenStddev", "StddevVoicedSegmentLengthSec"),
                 ("F0semitoneFrom27.5Hz_sma3nz__f0pause_meanSegLen", "MeanUnvoicedSegmentLength"),
                 ("F0semitoneFrom27.5Hz_sma3nz__f0pause_segLenStddev", "StddevUnvoicedSegmentLength")]


def _inst(spec, rows, names, period):
    return list(zip(fo.element_names(spec, names), fo.functionals(spec, rows, period)))


def egemaps_summary(levels, names, period=0.01):
    """levels / names: dicts keyed by the level names of the functionals' readers -> rows [n, K] as the functionals see them (first
    end-of-input tick) and their element names.  A multi-level reader sees min(rows) of its levels.  Returns (names, values[88])."""
    f0, ld = "gemapsv01b_lld_single_logF0_smo", "gemapsv01b_loudness_smo"
    z, nz = "egemapsv02_lldSetNoF0AndLoudnessZ_smo", "egemapsv02_lldSetNoF0AndLoudnessNz_smo"
    snz, sz, en = "egemapsv02_lldSetSpectralNz_smo", "egemapsv02_lldSetSpectralZ_smo", "egemapsv02_energyRMS"
    out = []
    out += _inst(F0_SPEC, levels[f0], names[f0], period)
    out += _inst(LOUD_SPEC, levels[ld], names[ld], period)
    out += _inst(MVR_SPEC, levels[z], names[z], period)
    n = min(len(levels[nz]), len(levels[snz]))                        # core/dataReader.cpp:375-380
    out += _inst(MVR_VOICED_SPEC, np.concatenate([levels[nz][:n], levels[snz][:n]], axis=1), list(names[nz]) + list(names[snz]), period)
    out += _inst(MEAN_UV_SPEC, levels[sz], names[sz], period)
    temporal = dict(_inst(PEAKS_SPEC, levels[ld], names[ld], period) + _inst(VOICED_SEG_SPEC, levels[f0], names[f0], period) +
                    _inst(PAUSE_SEG_SPEC, levels[f0], names[f0], period))
    out += [(new, temporal[old]) for old, new in TEMPORAL_PICK]
    leq = _inst(LEQ_SPEC, levels[en], names[en], period)
    out += [("equivalentSoundLevel_dBp", dbp(np.array([v for _, v in leq], F32))[0])]
    return [n_ for n_, _ in out], np.array([v for _, v in out], F32)
