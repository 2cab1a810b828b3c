"""CPU restatement of cFunctionals (frameMode = full) with the sub-components cFunctionalExtremes, cFunctionalMeans,
cFunctionalMoments, cFunctionalPercentiles, cFunctionalRegression (numpy, float64 accumulators like the reference) and
cFunctionalTimes, cFunctionalLpc, cFunctionalSegments (relTh / nonX / eqX), cFunctionalPeaks2 (float32 statement by statement:
their reference code works in FLOAT_DMEM).

TEST INFRASTRUCTURE ONLY (tests/, smoke, bench parity).  Pinned against the unmodified reference's -csvoutput / -arffoutput
rows: tests/golden/functionals_goldens.npz (scripts/make_golden_functionals.py), tests/test_functionals_cpu.py.

The second group is pinned by tests/golden/functionals_goldens2.npz (tests/configs/func_variants2.conf: the option sets of the
shipped ComParE_2016 / GeMAPS functionals blocks).

Citations relative to /root/reference/src/functionals.  One contour = one LLD column of one utterance:
  functionals.cpp:284-330  non-zero filter (nonZeroFuncts 1: != 0, 2: > 0), sorted copy, min / max / mean (double sum, divided
                           by the count, handed to the sub-components as float)
  functionals.cpp:215-256  element names <lld>_<value> (or <lld>__<functNameAppend>_<value>), values of element e contiguous
"""
import math

import numpy as np

SEGMENT, SECOND, FRAME = 0, 1, 2
F32 = np.float32


def _norm(own, own_set, master):
    """functionalComponent.hpp:67-76: the sub-component's own `norm` wins when it is set in the configuration, else the parent's
    masterTimeNorm when that is set, else the sub-component's default"""
    if own_set:
        return own
    return master if master is not None else own


class Spec:
    """mirror of include/osm_b200_functionals.h (field names = the reference's configuration fields)"""

    def __init__(self, enabled, non_zero=0, master_norm=None, name_append=None, extremes=None, means=None, moments=None,
                 percentiles=None, regression=None, times=None, lpc=None, segments=None, peaks2=None, onset=None, peaks=None, crossings=None, samples=None, dct=None):
        self.enabled, self.non_zero, self.master_norm, self.name_append = list(enabled), non_zero, master_norm, name_append
        self.extremes = dict(max=1, min=1, range=1, maxpos=1, minpos=1, amean=0, maxameandist=1, minameandist=1, norm=FRAME, norm_set=False)
        self.extremes.update(extremes or {})
        self.means = dict(amean=1, absmean=1, qmean=1, nzamean=1, nzabsmean=1, nzqmean=1, nzgmean=1, nnz=1, flatness=0, posamean=0,
                          negamean=0, posqmean=0, posrqmean=0, negqmean=0, negrqmean=0, rqmean=0, nzrqmean=0, norm=FRAME, norm_set=False)
        self.means.update(means or {})
        self.moments = dict(variance=1, stddev=1, skewness=1, kurtosis=1, amean=0, stddevNorm=0, doRatioLimit=0)
        self.moments.update(moments or {})
        self.percentiles = dict(quartile1=0, quartile2=0, quartile3=0, iqr12=0, iqr23=0, iqr13=0, percentile=[], pctlrange=[], interp=1)
        self.percentiles.update(percentiles or {})
        self.regression = dict(linregc1=1, linregc2=1, linregerrA=1, linregerrQ=1, qregc1=1, qregc2=1, qregc3=1, qregerrA=1, qregerrQ=1,
                               centroid=1, centroidNorm=SEGMENT, centroidUseAbsValues=1, normRegCoeff=0, normInputs=0, oldBuggyQerr=1,
                               centroidRatioLimit=1, doRatioLimit=0)
        self.regression.update(regression or {})
        self.times = dict(upleveltime25=1, downleveltime25=1, upleveltime50=1, downleveltime50=1, upleveltime75=1, downleveltime75=1,
                          upleveltime90=1, downleveltime90=1, risetime=1, falltime=1, leftctime=1, rightctime=1, duration=1,
                          buggySecNorm=1, norm=SEGMENT, norm_set=False)
        self.times.update(times or {})
        self.lpc = dict(lpGain=0, lpc=1, firstCoeff=0, order=5)
        self.lpc.update(lpc or {})
        self.segments = dict(maxNumSeg=20, segmentationAlgorithm="delta", thresholds=[0.0], X=0.0, XisRel=0, rangeRelThreshold=0.2,
                             numSegments=0, meanSegLen=0, maxSegLen=0, minSegLen=0, segLenStddev=0, segMinLng=3, segMinLng_set=False,
                             pauseMinLng=2, norm=SEGMENT, norm_set=False)
        self.segments.update(segments or {})
        self.peaks2 = {k: 0 for k in PEAKS2_NAMES}
        self.peaks2.update(dict(norm=FRAME, norm_set=False, relThresh=0.1, dynRelThresh=0, absThresh=None, doRatioLimit=1))
        self.peaks2.update(peaks2 or {})
        self.onset = dict(onsetPos=0, offsetPos=0, numOnsets=1, numOffsets=0, onsetRate=0, threshold=0.0, thresholdOnset=None, thresholdOffset=None,
                          useAbsVal=0, norm=SEGMENT, norm_set=False)
        self.onset.update(onset or {})
        self.peaks = dict(numPeaks=1, meanPeakDist=1, peakMean=1, peakMeanMeanDist=1, peakDistStddev=0, norm=FRAME, norm_set=False)
        self.peaks.update(peaks or {})
        self.crossings = dict(zcr=1, mcr=1, amean=0)
        self.crossings.update(crossings or {})
        self.samples = dict(samplepos=[i / 4.0 for i in range(5)])                  # functionalSamples.cpp:24,68-75
        self.samples.update(samples or {})
        self.dct = dict(firstCoeff=1, lastCoeff=6)                                  # functionalDCT.cpp:38-40
        self.dct.update(dct or {})


EXT_NAMES = ["max", "min", "range", "maxPos", "minPos", "amean", "maxameandist", "minameandist"]
EXT_KEYS = ["max", "min", "range", "maxpos", "minpos", "amean", "maxameandist", "minameandist"]
MEAN_NAMES = ["amean", "absmean", "qmean", "nzamean", "nzabsmean", "nzqmean", "nzgmean", "nnz", "flatness", "posamean", "negamean",
              "posqmean", "posrqmean", "negqmean", "negrqmean", "rqmean", "nzrqmean"]
MOM_NAMES = ["variance", "stddev", "skewness", "kurtosis", "amean"]
TIMES_NAMES = ["upleveltime25", "downleveltime25", "upleveltime50", "downleveltime50", "upleveltime75", "downleveltime75",
               "upleveltime90", "downleveltime90", "risetime", "falltime", "leftctime", "rightctime", "duration"]
SEG_NAMES = ["numSegments", "meanSegLen", "maxSegLen", "minSegLen", "segLenStddev"]
ONSET_NAMES = ["onsetPos", "offsetPos", "numOnsets", "numOffsets", "onsetRate"]                  # functionalOnset.cpp:29
PEAKS_NAMES = ["numPeaks", "meanPeakDist", "peakMean", "peakMeanMeanDist", "peakDistStddev"]     # functionalPeaks.cpp:29
CROSS_NAMES = ["zcr", "mcr", "amean"]                                                            # functionalCrossings.cpp:26
# functionalPeaks2.cpp:24-73: output order = index order of the FUNCT_* constants; the configuration field of a value is its
# name except for the three marked ones
PEAKS2_NAMES = ["numPeaks", "meanPeakDist", "meanPeakDistDelta", "peakDistStddev", "peakRangeAbs", "peakRangeRel", "peakMeanAbs",
                "peakMeanMeanDist", "peakMeanRel", "ptpAmpMeanAbs", "ptpAmpMeanRel", "ptpAmpStddevAbs", "ptpAmpStddevRel", "minRangeAbs",
                "minRangeRel", "minMeanAbs", "minMeanMeanDist", "minMeanRel", "mtmAmpMeanAbs", "mtmAmpMeanRel", "mtmAmpStddevAbs",
                "mtmAmpStddevRel", "meanRisingSlope", "maxRisingSlope", "minRisingSlope", "stddevRisingSlope", "meanFallingSlope",
                "maxFallingSlope", "minFallingSlope", "stddevFallingSlope", "covFallingSlope", "covRisingSlope"]
REG_NAMES = ["linregc1", "linregc2", "linregerrA", "linregerrQ", "qregc1", "qregc2", "qregc3", "qregerrA", "qregerrQ", "centroid"]


def value_names(spec):
    out = []
    for f in spec.enabled:
        if f == "Extremes":
            out += [n for n, k in zip(EXT_NAMES, EXT_KEYS) if spec.extremes[k]]
        elif f == "Means":
            out += [n for n in MEAN_NAMES if spec.means[n]]
        elif f == "Moments":
            out += [n for n in MOM_NAMES if spec.moments[n]]
            if spec.moments["stddevNorm"] == 2:
                out.append("stddevNorm")
            elif spec.moments["stddevNorm"] == 1:
                out.append("coeffOfVariation")
        elif f == "Percentiles":
            p = spec.percentiles
            out += [n for n, k in zip(["quartile1", "quartile2", "quartile3", "iqr1-2", "iqr2-3", "iqr1-3"],
                                      ["quartile1", "quartile2", "quartile3", "iqr12", "iqr23", "iqr13"]) if p[k]]
            out += ["percentile%.1f" % (100.0 * x) for x in p["percentile"]]
            out += ["pctlrange%d-%d" % tuple(r) for r in p["pctlrange"]]
        elif f == "Regression":
            out += [n for n in REG_NAMES if spec.regression[n]]
        elif f == "Times":
            out += [n for n in TIMES_NAMES if spec.times[n]]
        elif f == "Lpc":                                           # functionalLpc.cpp:84-96
            l = spec.lpc
            out += (["lpgain"] if l["lpGain"] else []) + (["lpc%d" % i for i in range(l["firstCoeff"], l["order"])] if l["lpc"] else [])
        elif f == "Segments":
            out += [n for n in SEG_NAMES if spec.segments[n]]
        elif f == "Peaks2":
            out += [n for n in PEAKS2_NAMES if spec.peaks2[n]]
        elif f == "Onset":
            out += [n for n in ONSET_NAMES if spec.onset[n]]
        elif f == "Peaks":
            out += [n for n in PEAKS_NAMES if spec.peaks[n]]
        elif f == "Crossings":
            out += [n for n in CROSS_NAMES if spec.crossings[n]]
        elif f == "Samples":                                       # functionalSamples.cpp:89-95
            out += ["samples%.3f" % x for x in spec.samples["samplepos"]]
        elif f == "DCT":                                           # functionalDCT.cpp:103-110
            out += ["dct%d" % k for k in range(spec.dct["firstCoeff"], spec.dct["lastCoeff"] + 1)]
        else:
            raise ValueError(f)
    return out


def element_names(spec, lld_names):
    vn = value_names(spec)
    if spec.name_append:
        return ["%s__%s_%s" % (l, spec.name_append, v) for l in lld_names for v in vn]
    return ["%s_%s" % (l, v) for l in lld_names for v in vn]


def _interp_pctl(p, s):
    """functionalPercentiles.cpp:317-336: linear interpolation between the neighbours in the sorted contour, float products"""
    N = len(s)
    idx = p * float(N - 1)
    i1, i2 = min(max(int(math.floor(idx)), 0), N - 1), min(max(int(math.ceil(idx)), 0), N - 1)
    if i1 != i2:
        w1, w2 = idx - float(i1), float(i2) - idx
        return F32(F32(s[i1] * F32(w2)) + F32(s[i2] * F32(w1)))
    return F32(s[i1])


def contour(spec, x, period):
    """all enabled values of one contour x (float32 [T]) -> float32 list, in output order"""
    x = np.asarray(x, np.float32)
    if spec.non_zero == 2:
        x = x[x > 0]
    elif spec.non_zero:
        x = x[x != 0]
    nvals = len(value_names(spec))
    N = len(x)
    if N == 0:
        return [F32(0)] * nvals                                   # every sub-component returns 0 values -> zero fill (:316-320)
    xd = x.astype(np.float64)
    mn, mx = F32(x.min()), F32(x.max())
    mean = F32(xd.sum() / float(N))                                 # double sum / count, passed on as float (:300-306,:312)
    out = []
    for f in spec.enabled:
        if f == "Extremes":                                        # functionalExtremes.cpp:89-132
            e = spec.extremes
            maxpos, minpos = F32(int(np.argmax(x == mx))), F32(int(np.argmax(x == mn)))
            nrm = _norm(e["norm"], e["norm_set"], spec.master_norm)
            if nrm == SEGMENT:
                maxpos, minpos = F32(maxpos / F32(N)), F32(minpos / F32(N))
            elif nrm == SECOND and F32(period) != 0:
                maxpos, minpos = F32(maxpos * F32(period)), F32(minpos * F32(period))
            vals = dict(max=mx, min=mn, range=F32(mx - mn), maxpos=maxpos, minpos=minpos, amean=mean, maxameandist=F32(mx - mean),
                        minameandist=F32(mean - mn))
            out += [vals[k] for k in EXT_KEYS if e[k]]
        elif f == "Means":                                         # functionalMeans.cpp:104-262
            m = spec.means
            fa = np.abs(xd)
            nz = xd != 0
            nnz = int(nz.sum())
            absmean, qmean = fa.sum() / N, (xd * xd).sum() / N
            nzamean = nzabsmean = nzqmean = nzgmean = 0.0
            if nnz > 0:
                nzamean, nzabsmean, nzqmean = xd[nz].sum() / nnz, fa[nz].sum() / nnz, (xd[nz] ** 2).sum() / nnz
                nzgmean = math.exp(np.log(fa[nz]).sum() / nnz)
            pos, neg = xd[xd > 0], xd[xd < 0]
            posamean, posqmean = (pos.sum() / len(pos), (pos ** 2).sum() / len(pos)) if len(pos) else (0.0, 0.0)
            negamean, negqmean = (neg.sum() / len(neg), (neg ** 2).sum() / len(neg)) if len(neg) else (0.0, 0.0)
            nrm = _norm(m["norm"], m["norm_set"], spec.master_norm)
            nnzv = F32(nnz) if nrm == FRAME else (F32(F32(nnz) / F32(N)) if nrm == SEGMENT else F32(F32(nnz) / F32(period)))
            vals = dict(amean=mean, absmean=F32(absmean), qmean=F32(qmean), nzamean=F32(nzamean), nzabsmean=F32(nzabsmean),
                        nzqmean=F32(nzqmean), nzgmean=F32(nzgmean), nnz=nnzv, flatness=F32(nzgmean / absmean) if absmean != 0 else F32(1),
                        posamean=F32(posamean), negamean=F32(negamean), posqmean=F32(posqmean), posrqmean=F32(math.sqrt(posqmean)),
                        negqmean=F32(negqmean), negrqmean=F32(math.sqrt(negqmean)), rqmean=F32(math.sqrt(qmean)), nzrqmean=F32(math.sqrt(nzqmean)))
            out += [vals[k] for k in MEAN_NAMES if m[k]]
        elif f == "Moments":                                       # functionalMoments.cpp:89-168
            m = spec.moments
            d = xd - float(mean)
            m2, m3, m4 = (d * d).sum() / N, (d ** 3).sum(), (d ** 4).sum()
            sq = math.sqrt(m2)
            if m["variance"]:
                out.append(F32(m2))
            if m["stddev"]:
                out.append(F32(sq) if m2 > 0 else F32(0))
            if m["skewness"]:
                out.append(F32(m3 / (N * m2 * sq)) if m2 > 0 else F32(0))
            if m["kurtosis"]:
                out.append(F32(m4 / (N * m2 * m2)) if m2 > 0 else F32(0))
            if m["amean"]:
                out.append(mean)
            if m["stddevNorm"]:
                if m2 > 0:
                    ml = float(abs(mean)) if m["stddevNorm"] == 1 else float(mean)
                    if m["doRatioLimit"]:                          # functionalMoments.cpp:144-151
                        out.append(_ratio_limit(F32(sq / ml), 10.0, 20.0) if ml != 0 else F32(20.0))
                    else:
                        out.append(F32(sq / (ml if ml != 0 else 1.0)))
                else:
                    out.append(F32(0))
        elif f == "Percentiles":                                   # functionalPercentiles.cpp:338-430
            p = spec.percentiles
            s = np.sort(x)
            get = (lambda q: _interp_pctl(q, s)) if p["interp"] else (lambda q: F32(s[min(max(int(math.floor(q * (N - 1) + 0.5)), 0), N - 1)]))   # C round(): half away from zero
            q1, q2, q3 = get(0.25), get(0.50), get(0.75)
            vals = dict(quartile1=q1, quartile2=q2, quartile3=q3, iqr12=F32(q2 - q1), iqr23=F32(q3 - q2), iqr13=F32(q3 - q1))
            out += [vals[k] for k in ["quartile1", "quartile2", "quartile3", "iqr12", "iqr23", "iqr13"] if p[k]]
            pv = [get(q) for q in p["percentile"]]
            out += pv
            out += [F32(abs(F32(pv[b] - pv[a]))) for a, b in p["pctlrange"]]
        elif f == "Regression":                                    # functionalRegression.cpp:141-428
            r = spec.regression
            Nd = float(N)
            rng = float(F32(mx - mn))                                # FLOAT_DMEM expression
            rinv = 1.0 / rng if rng > 0 else 0.0
            if rng <= 0:
                rng = 1.0                                            # :151-157
            ii = np.arange(N, dtype=np.float64)
            num, num2 = (xd * ii).sum(), (xd * ii * ii).sum()
            asum = float(mean) * Nd
            if r["centroidUseAbsValues"]:
                asa = np.abs(xd).sum()
                centroid = (np.abs(xd) * ii).sum() / asa if asa != 0 else 0.0
            else:
                centroid = num / asum if asum != 0 else 0.0
            if r["centroidRatioLimit"]:                              # :206-209, before the time normalisation
                centroid = float(_ratio_limit(F32(centroid), float(F32(Nd)), float(F32(Nd))))
            if r["centroidNorm"] == SECOND:
                centroid *= period
            elif r["centroidNorm"] == SEGMENT:
                centroid /= Nd
            enq = any(r[k] for k in ("qregc1", "qregc2", "qregc3", "qregerrA", "qregerrQ", "centroid"))
            a = b = c = 0.0
            if N > 1:
                nnm1 = Nd * (Nd - 1.0)
                S1, S2 = nnm1 / 2.0, nnm1 * (2.0 * Nd - 1.0) / 6.0
                s1d = S1 / S2
                tmp = Nd - S1 * s1d
                t = 0.0 if tmp == 0 else (asum - num * s1d) / tmp
                m = (num - t * S1) / S2
                S3 = S1 * S1
                n1 = Nd - 1.0
                S4 = S2 * (3.0 * (n1 * n1 + n1) - 1.0) / 5.0
                if enq:
                    det = S4 * S2 * Nd + 2.0 * S3 * S1 * S2 - S2 * S2 * S2 - S3 * S3 * Nd - S3 * S4
                    if det != 0:
                        a = ((S2 * Nd - S3) * num2 + (S1 * S2 - S3 * Nd) * num + (S3 * S1 - S2 * S2) * asum) / det
                        b = ((S1 * S2 - S3 * Nd) * num2 + (S4 * Nd - S2 * S2) * num + (S3 * S2 - S4 * S1) * asum) / det
                        c = ((S3 * S1 - S2 * S2) * num2 + (S3 * S2 - S4 * S1) * num + (S4 * S2 - S3 * S3) * asum) / det
            else:
                m, t, c = 0.0, float(x[0]), float(x[0])
            e = xd - (m * ii + t)
            if r["normInputs"]:
                e = e * rinv
            lea, leq = np.abs(e).sum(), (e * e).sum()
            qea = qeq = 0.0
            if enq:
                e = xd - (a * ii * ii + b * ii + c)
                if r["normInputs"]:
                    e = e * rinv
                qea, qeq = np.abs(e).sum(), (e * e).sum()
            if r["doRatioLimit"]:                                    # :328-335
                l1 = float(F32(rng / 10.0))
                m = float(_ratio_limit(F32(m), l1, float(F32(rng / 10.0 + 0.01))))
                a = float(_ratio_limit(F32(a), float(F32(math.sqrt(rng / 10.0))), float(F32(math.sqrt(rng / 10.0) + 0.01))))
                b = float(_ratio_limit(F32(b), l1, float(F32(rng / 10.0 + 0.01))))
            if r["normRegCoeff"] == 1:
                m *= Nd - 1.0; a *= (Nd - 1.0) ** 2; b *= Nd - 1.0
            elif r["normRegCoeff"] == 2:
                one = 1.0 / period
                m *= one; a *= one * one; b *= one
            if r["normInputs"]:
                m *= rinv; t = (t - float(mn)) * rinv; a *= rinv; b *= rinv; c = (c - float(mn)) * rinv
            fin = lambda v: v if math.isfinite(v) else 0.0
            vals = dict(linregc1=fin(m), linregc2=fin(t), linregerrA=fin(lea / Nd), linregerrQ=fin(leq / Nd), qregc1=fin(a), qregc2=fin(b),
                        qregc3=fin(c), qregerrA=fin(qea) if r["oldBuggyQerr"] else fin(qea / Nd),
                        qregerrQ=fin(qeq) if r["oldBuggyQerr"] else fin(qeq / Nd), centroid=fin(centroid))
            out += [F32(vals[k]) for k in REG_NAMES if r[k]]
        elif f == "Times":
            out += _times(spec, x, mn, mx, period)
        elif f == "Lpc":
            out += _lpc(spec, x)
        elif f == "Segments":
            out += _segments(spec, x, mn, mx, mean, period)
        elif f == "Peaks2":
            out += _peaks2(spec, x, mn, mx, mean, period)
        elif f == "Onset":
            out += _onset(spec, x, period)
        elif f == "Peaks":
            out += _peaks_old(spec, x, period)
        elif f == "Crossings":
            out += _crossings(spec, x)
        elif f == "Samples":                                       # functionalSamples.cpp:99-116
            out += [x[int((float(F32(N)) - 1.0) * float(pos))] for pos in spec.samples["samplepos"]]
        elif f == "DCT":
            out += _dct(spec, x)
    assert len(out) == nvals
    return out


def _onset(spec, x, period):
    """functionalOnset.cpp:95-153: a two-state machine over the contour (thresholdOnset / thresholdOffset, :77-81)"""
    o = spec.onset
    thr_on = F32(o["threshold"] if o["thresholdOnset"] is None else o["thresholdOnset"])
    thr_off = F32(o["threshold"] if o["thresholdOffset"] is None else o["thresholdOffset"])
    N = len(x)
    onset_pos = offset_pos = -1
    n_on = n_off = 0
    oo = 1 if x[0] > thr_on else 0
    for i in range(1, N):
        cur = F32(abs(x[i])) if o["useAbsVal"] else x[i]
        if cur > thr_on and oo == 0:
            n_on += 1
            if onset_pos == -1:
                onset_pos = i
            oo = 1
        if cur <= thr_off and oo == 1:
            n_off += 1
            offset_pos = i
            oo = 0
    if offset_pos == -1:
        offset_pos = N - 1
    if onset_pos == -1:
        onset_pos = 0
    nrm = _norm(o["norm"], o["norm_set"], spec.master_norm)
    T = F32(period)
    if nrm == SEGMENT:
        pos = [F32(F32(onset_pos) / F32(N)), F32(F32(offset_pos) / F32(N))]
    elif nrm == SECOND:
        pos = [F32(F32(onset_pos) * T), F32(F32(offset_pos) * T)]
    else:
        pos = [F32(onset_pos), F32(offset_pos)]
    vals = dict(onsetPos=pos[0], offsetPos=pos[1], numOnsets=F32(n_on), numOffsets=F32(n_off), onsetRate=F32(F32(n_on) / F32(F32(N) * T)))
    return [vals[k] for k in ONSET_NAMES if o[k]]


def _peaks_old(spec, x, period):
    """functionalPeaks.cpp:96-213 with overlapFlag = 1 (the default): float running sums, the 0.11 / 0.09 range tests in double"""
    k = spec.peaks
    N = len(x)
    mean = x[0]
    for i in range(1, N):
        mean = F32(mean + x[i])
    mean = F32(mean / F32(N))
    rng = F32(x.max() - x.min())
    peak_dist, peak_mean, last_min, last_max = F32(0), F32(0), F32(0), F32(0)
    dists = []
    n_peaks, curmax, lastmax_pos, flag = 0, 0, -1, 0
    llv, lv = x[0], (x[1] if N > 1 else F32(0))
    for i in range(2, N):
        if llv < lv and lv > x[i]:
            if not flag:
                last_max = x[i]
            elif x[i] > last_max:
                last_max, curmax = x[i], i
            if float(F32(last_max - last_min)) > 0.11 * float(rng):
                flag, curmax = 1, i
        elif llv > lv and lv < x[i]:
            last_min = x[i]
        if flag and (float(x[i]) < float(last_max) - 0.09 * float(rng) or i == N - 1):
            n_peaks += 1
            peak_mean = F32(peak_mean + last_max)
            if lastmax_pos >= 0:
                d = F32(curmax - lastmax_pos)
                peak_dist = F32(peak_dist + d)
                dists.append(int(d))
            lastmax_pos, flag = curmax, 0
        llv, lv = lv, x[i]
    stddev = F32(0)
    if dists:
        peak_dist = F32(peak_dist / F32(len(dists)))
        for d in dists:
            t = F32(F32(d) - peak_dist)
            stddev = F32(stddev + F32(t * t))
        stddev = F32(np.sqrt(F32(stddev / F32(len(dists)))))
    else:
        peak_dist = F32(N + 1)
    nrm = _norm(k["norm"], k["norm_set"], spec.master_norm)
    if nrm == SECOND:
        peak_dist, stddev = F32(peak_dist * F32(period)), F32(stddev * F32(period))
    elif nrm == SEGMENT:
        peak_dist, stddev = F32(peak_dist / F32(N)), F32(stddev / F32(N))
    peak_mean = F32(peak_mean / F32(n_peaks)) if n_peaks > 0 else F32(0)
    vals = dict(numPeaks=F32(n_peaks), meanPeakDist=peak_dist, peakMean=peak_mean, peakMeanMeanDist=F32(peak_mean - mean), peakDistStddev=stddev)
    return [vals[n] for n in PEAKS_NAMES if k[n]]


def _crossings(spec, x):
    """functionalCrossings.cpp:64-97: zero crossings on float products, mean crossings on double differences"""
    c = spec.crossings
    N = len(x)
    amean = 0.0
    if c["mcr"] or c["amean"]:
        amean = float(x[0])
        for i in range(1, N):
            amean += float(x[i])
        amean /= float(N)
    zcr = mcr = 0
    for i in range(1, N - 1):
        a, b, d = x[i - 1], x[i], x[i + 1]
        if (F32(a * d) <= 0 and b == 0) or F32(a * b) < 0:
            zcr += 1
        if c["mcr"]:
            am, bm, dm = float(a) - amean, float(b) - amean, float(d) - amean
            if (am * dm <= 0.0 and bm == 0.0) or am * bm < 0.0:
                mcr += 1
    vals = dict(zcr=F32(zcr / float(N)), mcr=F32(mcr / float(N)), amean=F32(amean))
    return [vals[n] for n in CROSS_NAMES if c[n]]


def _dct(spec, x):
    """functionalDCT.cpp:85-135: table (float)cos(pi * i / N * ((float)m + 0.5)), float sum over the contour, times (float)sqrt(2 / N)"""
    N = len(x)
    m = np.arange(N, dtype=np.float32).astype(np.float64) + 0.5
    out = []
    for i in range(spec.dct["firstCoeff"], spec.dct["lastCoeff"] + 1):
        tab = np.cos(math.pi * float(i) / float(N) * m).astype(np.float32)
        acc = F32(0)
        for v in (x * tab).astype(np.float32):                     # float product, float running sum
            acc = F32(acc + v)
        acc = F32(acc * F32(math.sqrt(2.0 / float(N))))
        out.append(acc if np.isfinite(acc) else F32(0))
    return out


def _times(spec, x, mn, mx, period):
    """functionalTimes.cpp:245-371 (upleveltime[] / downleveltime[] arrays and useRobustPercentileRange are not restated)"""
    t = spec.times
    N = len(x)
    Nind = F32(N)
    Norm, Norm1, Norm2 = Nind, F32(Nind - F32(1)), F32(Nind - F32(2))
    nrm = _norm(t["norm"], t["norm_set"], spec.master_norm)
    T = F32(1)
    if nrm == SECOND:
        T = F32(period)
        if T != 0:
            if t["buggySecNorm"]:
                Norm, Norm1, Norm2 = F32(Norm / T), F32(Norm1 / T), F32(Norm2 / T)
            else:
                Norm = F32(F32(1.0) / T)
                Norm1 = F32(Norm1 / F32(Nind * T))
                Norm2 = F32(Norm2 / F32(Nind * T))
    if nrm == FRAME:
        Norm, Norm1, Norm2 = F32(1), F32(Norm1 / Nind), F32(Norm2 / Nind)
    rng = F32(mx - mn)
    lv = [F32(F32(F32(q) * rng) + mn) for q in (0.25, 0.50, 0.75, 0.90)]
    n25, n50, n75, n90 = [int((x <= l).sum()) for l in lv]
    nR, nF = int((x[:-1] < x[1:]).sum()), int((x[:-1] > x[1:]).sum())
    a1, a2 = (x[1:-1] - x[:-2]).astype(np.float32), (x[2:] - x[1:-1]).astype(np.float32)
    nRC, nLC = int((a2 < a1).sum()), int((a1 < a2).sum())
    v = {}
    for nm, cnt in (("25", n25), ("50", n50), ("75", n75), ("90", n90)):
        v["upleveltime" + nm] = F32(F32(N - cnt) / Norm)
        v["downleveltime" + nm] = F32(F32(cnt) / Norm)
    v["risetime"] = F32(F32(nR) / Norm1) if Norm1 != 0 else F32(0)
    v["falltime"] = F32(F32(nF) / Norm1) if Norm1 != 0 else F32(0)
    v["leftctime"] = F32(F32(nLC) / Norm2) if Norm2 != 0 else F32(0)
    v["rightctime"] = F32(F32(nRC) / Norm2) if Norm2 != 0 else F32(0)
    v["duration"] = F32(F32(N) * T) if nrm == SECOND else F32(N)
    return [v[k] for k in TIMES_NAMES if t[k]]


def _durbin(r, p):
    """smileDsp_calcLpcAcf (smileutil/smileUtil.c:1572-1627), float: predictor coefficients a[0..p-1] and the final error"""
    a = [F32(0)] * (p + 1)
    if r[0] == 0:
        return a[:p], F32(0)
    e = F32(r[0])
    for m in range(1, p + 1):
        s = F32(F32(1.0) * r[m])
        for i in range(1, m):
            s = F32(s + F32(a[i - 1] * r[m - i]))
        km = F32(F32(F32(-1.0) / e) * s)
        a[m - 1] = km
        for i in range(1, m // 2 + 1):
            xx = a[i - 1]
            a[i - 1] = F32(a[i - 1] + F32(km * a[m - i - 1]))
            if i < m // 2 or (m & 1) == 1:
                a[m - i - 1] = F32(a[m - i - 1] + F32(km * xx))
        e = F32(e * F32(F32(1.0) - F32(km * km)))
        if e == 0:
            for i in range(m, p + 1):
                a[i] = F32(0)
            break
    return a[:p], e


def _lpc(spec, x):
    """functionalLpc.cpp:98-125: smileDsp_autoCorr (smileUtil.c:1560-1569, sequential float sums) + Durbin on the contour"""
    l = spec.lpc
    order, N = l["order"], len(x)
    acf = []
    for lag in range(order + 1):
        acc = F32(0)
        for i in range(lag, N):
            acc = F32(acc + F32(x[i] * x[i - lag]))
        acf.append(acc)
    a, gain = _durbin(acf, order)
    out = []
    if l["lpGain"]:
        out.append(F32(gain / F32(N)))
    if l["lpc"]:
        out += [F32(a[i]) for i in range(l["firstCoeff"], order)]
    return out


def _segments(spec, x, mn, mx, mean, period):
    """functionalSegments.cpp: relTh (process_SegThresh :305-367), nonX (:658-726), eqX (:729-797), statistics + output :800-960"""
    g = spec.segments
    N = len(x)
    maxNumSeg = g["maxNumSeg"]
    seglens = []
    st = dict(mean=0, mx=0, mn=0)

    def add(i, last):                                   # addNewSegment :240-262
        L = i - last
        if len(seglens) < maxNumSeg:
            st["mean"] += L
            seglens.append(L)
            if L > st["mx"]:
                st["mx"] = L
            if st["mn"] == 0 or L < st["mn"]:
                st["mn"] = L
        return i

    rng = F32(mx - mn)
    alg = g["segmentationAlgorithm"]
    if alg == "relTh":
        th = [F32(mn + F32(rng * F32(t))) for t in g["thresholds"]]
        segMinLng = max(g["segMinLng"], 1)
        if not g["segMinLng_set"]:
            segMinLng = max(N // maxNumSeg - 1, 2)
        ravgLng = 3                                     # :333 (the ravgLng option is not used by this method)
        lastSeg = int(-segMinLng / 2)                   # C integer division truncates toward zero
        ravg, raLast = F32(0), F32(0)
        for i in range(N):
            ravg = F32(ravg + x[i])
            if i >= ravgLng:
                ravg = F32(ravg - x[i - ravgLng])
            ra = F32(ravg / F32(min(i + 1, ravgLng)))
            cross = any((ra > t and raLast <= t) or (ra < t and raLast >= t) for t in th)
            raLast = ra
            if cross and i - lastSeg > segMinLng:
                lastSeg = add(i, lastSeg)
    elif alg == "NArelTh":                              # process_SegThreshNoavg :369-413: the samples themselves, from frame 1
        th = [F32(mn + F32(rng * F32(t))) for t in g["thresholds"]]
        segMinLng = max(g["segMinLng"], 1)
        if not g["segMinLng_set"]:
            segMinLng = max(N // maxNumSeg - 1, 2)
        lastSeg = int(-segMinLng / 2)
        for i in range(1, N):
            cross = any((x[i] > t and x[i - 1] <= t) or (x[i] < t and x[i - 1] >= t) for t in th)
            if cross and i - lastSeg > segMinLng:
                lastSeg = add(i, lastSeg)
    elif alg in ("nonX", "eqX"):
        X = F32(mn + F32(rng * F32(g["X"]))) if g["XisRel"] else F32(g["X"])
        segMinLng, pauseMinLng = max(g["segMinLng"], 1), max(g["pauseMinLng"], 1)
        inSeg = segStart = segEnd = 0
        startIdx = 0
        for i in range(N):
            hit = (x[i] != X) if alg == "nonX" else (x[i] == X)
            if hit:
                if inSeg == 1:
                    segEnd = 0
                    segStart += 1
                    if segStart >= segMinLng:
                        segStart = 0
                        inSeg = 2
                elif inSeg == 0:
                    segStart += 1
                    startIdx = i
                    inSeg = 1
                else:
                    segEnd = 0
            else:
                if inSeg == 2:
                    segStart = 0
                    segEnd += 1
                    if segEnd >= pauseMinLng:
                        inSeg = 0
                        add(i - segEnd, startIdx)
                        segEnd = 0
                elif inSeg == 1:
                    segEnd += 1
                    if segEnd >= pauseMinLng:
                        inSeg = segEnd = segStart = 0
        if inSeg == 2:
            segEnd += 1
            add(N - segEnd, startIdx)
    else:
        raise ValueError(alg)
    nS = len(seglens)
    m = F32(F32(st["mean"]) / F32(nS)) if nS > 1 else F32(st["mean"])
    dev = F32(0)
    for L in seglens:
        d = F32(F32(L) - m)
        dev = F32(dev + F32(d * d))
    dev = F32(math.sqrt(F32(dev / F32(nS)))) if nS > 1 else F32(0)      # sqrt(double(float)) -> float
    nrm = _norm(g["norm"], g["norm_set"], spec.master_norm)
    T = F32(period)
    Tn = T if T != 0 else F32(1)
    v = {}
    if nrm == SECOND:
        v["numSegments"] = F32(F32(nS) / F32(Tn * F32(N)))
        v["meanSegLen"], v["maxSegLen"], v["minSegLen"], v["segLenStddev"] = F32(m * Tn), F32(F32(st["mx"]) * Tn), F32(F32(st["mn"]) * Tn), F32(dev * Tn)
    elif nrm == SEGMENT:
        v["numSegments"] = F32(F32(nS) / F32(maxNumSeg))
        v["meanSegLen"], v["maxSegLen"], v["minSegLen"], v["segLenStddev"] = F32(m / F32(N)), F32(F32(st["mx"]) / F32(N)), F32(F32(st["mn"]) / F32(N)), F32(dev / F32(N))
    else:
        v["numSegments"] = F32(nS)
        v["meanSegLen"], v["maxSegLen"], v["minSegLen"], v["segLenStddev"] = m, F32(st["mx"]), F32(st["mn"]), dev
    return [v[k] for k in SEG_NAMES if g[k]]


def _ratio_limit(x, lim1=10.0, lim2=10.0):
    """smileMath_ratioLimit (smileutil/smileUtil.c:602-614) with smileMath_tanh / smileMath_logistic (:590-600): the argument of
    tanh is a FLOAT_DMEM, the logistic is evaluated in double and returned as FLOAT_DMEM"""
    def tanh_(a):
        a = F32(a)
        z = F32(F32(2.0) * a)
        lim = F32(math.log(float(np.finfo(np.float32).max)))
        lg = F32(1.0) if z > lim else (F32(0.0) if z < -lim else F32(1.0 / (1.0 + math.exp(-float(z)))))
        return F32(F32(F32(2.0) * lg) - F32(1.0))
    x = F32(x)
    l1, l2 = F32(lim1), F32(lim2)
    if x > l1:
        return F32(F32(tanh_((math.sqrt(float(x) - float(l1) + 1.0) - 1.0) / (float(l2) * 0.5)) * l2) + l1)
    if x < -l1:
        return F32(F32(tanh_(-(math.sqrt(-1.0 * (float(x) + float(l1)) + 1.0) - 1.0) / (float(l2) * 0.5)) * l2) - l1)
    return x


def _peaks2(spec, x, mn, mx, mean, period):
    """functionalPeaks2.cpp:296-915.  The reference's doubly linked list of extrema becomes a Python list of [type, x, y] with
    removal by identity; the statement order of the three pruning passes and of the statistics is kept, FLOAT_DMEM = float32."""
    c = spec.peaks2
    N = len(x)
    rng = F32(mx - mn)
    absT = F32(c["absThresh"]) if c["absThresh"] is not None else F32(F32(c["relThresh"]) * rng)
    dyn = c["dynRelThresh"] and c["absThresh"] is None
    relT = F32(c["relThresh"])

    def below(diff, base):                              # isBelowThresh :270-294
        if dyn:
            if base == 0:
                return diff != 0
            return abs(float(F32(diff / base))) < relT  # fabs of a float quotient, compared with the float threshold
        return diff < absT

    lst = []
    for i in range(2, N - 2):                           # step 1 :320-327
        if x[i] > x[i - 1] and x[i] > x[i + 1]:
            lst.append([1, i, F32(x[i])])
        elif x[i] < x[i - 1] and x[i] < x[i + 1]:
            lst.append([0, i, F32(x[i])])

    def remove(el):
        for k, e in enumerate(lst):
            if e is el:
                del lst[k]
                return

    # step 2a :330-392
    lastVal = lastMin = lastMax = F32(x[0])
    maxFlag = minFlag = 0
    lastMaxPtr = None
    k = 0
    while k < len(lst):
        el = lst[k]
        nxt = lst[k + 1] if k + 1 < len(lst) else None
        if el[0] == 1:
            if below(F32(abs(F32(el[2] - lastVal))), min(el[2], lastVal)):
                if below(F32(el[2] - lastMin), lastMin):
                    remove(el)
                else:
                    if float(el[2]) > float(lastMax) * 1.05:        # FLOAT_DMEM * double literal: compared in double
                        if lastMaxPtr is not None:
                            remove(lastMaxPtr)
                        lastMax = el[2]
                        lastMaxPtr = el
                    else:
                        if minFlag:
                            lastMax = el[2]
                            lastMaxPtr = el
                        else:
                            remove(el)
                    maxFlag, minFlag = 1, 0
            else:
                maxFlag, minFlag = 1, 0
                lastMax = el[2]
                lastMaxPtr = el
        else:
            if not below(F32(abs(F32(el[2] - lastVal))), min(el[2], lastVal)):
                minFlag, maxFlag = 1, 0
                lastMin = el[2]
        lastVal = el[2]
        k = next((q for q, e in enumerate(lst) if e is nxt), len(lst)) if nxt is not None else len(lst)
    # step 2b :395-412
    lastMax = F32(x[0])
    for el in list(lst):
        if el[0] == 0:
            if below(F32(lastMax - el[2]), el[2]):
                remove(el)
        else:
            lastMax = el[2]
    # step 3 :415-466
    lastMax = lastMin = F32(x[0])
    minFlag = 0
    init = 1
    lastMinPtr = lastMaxPtr = None
    for el in list(lst):
        if not any(e is el for e in lst):
            continue
        if el[0] == 0:
            if not minFlag or init:
                lastMin, lastMinPtr, minFlag, init = el[2], el, 1, 0
            else:
                if el[2] >= lastMin:
                    remove(el)
                else:
                    if lastMinPtr is not el:
                        remove(lastMinPtr)
                        lastMinPtr, lastMin = el, el[2]
        else:
            if minFlag or init:
                lastMax, lastMaxPtr, minFlag, init = el[2], el, 0, 0
            else:
                if el[2] <= lastMax:
                    remove(el)
                else:
                    if lastMaxPtr is not el:
                        remove(lastMaxPtr)
                        lastMaxPtr, lastMax = el, el[2]
    # statistics, first pass :470-545
    Z = F32(0)
    peakMax = peakMin = peakDist = peakDiff = peakMean = Z
    minMax = minMin = minDist = minDiff = minMean = Z
    nPeakDist = nPeaks = nMinDist = nMins = 0
    lastMaxPtr = lastMinPtr = None
    for el in lst:
        if el[0] == 0:
            if lastMinPtr is None:
                lastMinPtr, minMin, minMax = el, el[2], el[2]
            else:
                nMinDist += 1
                minDist = F32(minDist + F32(el[1] - lastMinPtr[1]))
                minDiff = F32(minDiff + F32(abs(F32(el[2] - lastMinPtr[2]))))
                minMin, minMax = min(minMin, el[2]), max(minMax, el[2])
                lastMinPtr = el
            minMean = F32(minMean + el[2])
            nMins += 1
        else:
            if lastMaxPtr is None:
                lastMaxPtr, peakMin, peakMax = el, el[2], el[2]
            else:
                nPeakDist += 1
                peakDist = F32(peakDist + F32(el[1] - lastMaxPtr[1]))
                peakDiff = F32(peakDiff + F32(abs(F32(el[2] - lastMaxPtr[2]))))
                peakMin, peakMax = min(peakMin, el[2]), max(peakMax, el[2])
                lastMaxPtr = el
            peakMean = F32(peakMean + el[2])
            nPeaks += 1
    if nPeaks > 1:                                      # :548-561 (sic: a single peak keeps its sum, minima divide from one on)
        peakMean = F32(peakMean / F32(nPeaks))
        if nPeakDist > 1:
            peakDist, peakDiff = F32(peakDist / F32(nPeakDist)), F32(peakDiff / F32(nPeakDist))
    if nMins > 0:
        minMean = F32(minMean / F32(nMins))
        if nMinDist > 1:
            minDist, minDiff = F32(minDist / F32(nMinDist)), F32(minDiff / F32(nMinDist))
    # second pass :564-594 (sic: the peak deviations are taken against the last MINIMUM)
    peakSdDist = peakSdDiff = minSdDist = minSdDiff = Z
    lastMaxPtr = lastMinPtr = None
    for el in lst:
        if el[0] == 0:
            if lastMinPtr is None:
                lastMinPtr = el
            else:
                d = F32(F32(el[1] - lastMinPtr[1]) - minDist)
                minSdDist = F32(minSdDist + F32(d * d))
                d = F32(F32(abs(F32(el[2] - lastMinPtr[2]))) - minDiff)
                minSdDiff = F32(minSdDiff + F32(d * d))
                lastMinPtr = el
        else:
            if lastMaxPtr is None:
                lastMaxPtr = el
            else:
                d = F32(F32(el[1] - lastMinPtr[1]) - peakDist)
                peakSdDist = F32(peakSdDist + F32(d * d))
                d = F32(F32(abs(F32(el[2] - lastMinPtr[2]))) - peakDiff)
                peakSdDiff = F32(peakSdDiff + F32(d * d))
                lastMaxPtr = el
    sq = lambda v: F32(math.sqrt(float(v))) if v > 0 else F32(0)
    if nPeakDist > 1:
        peakSdDist, peakSdDiff = F32(peakSdDist / F32(nPeakDist)), F32(peakSdDiff / F32(nPeakDist))
    peakSdDist, peakSdDiff = sq(peakSdDist), sq(peakSdDiff)
    if nMinDist > 1:
        minSdDist, minSdDiff = F32(minSdDist / F32(nMinDist)), F32(minSdDiff / F32(nMinDist))
    minSdDist, minSdDiff = sq(minSdDist), sq(minSdDiff)
    # slopes :610-730
    meanRise = meanFall = minRise = maxRise = minFall = maxFall = sdRise = sdFall = Z
    nRising = nFalling = 0
    enabSlope = any(c[k] for k in PEAKS2_NAMES[22:])
    if enabSlope:
        T = F32(period)
        lastIsMax = -1
        lastMax = lastMin = F32(x[0])
        lastMaxPos = lastMinPos = 0

        def rise(s):
            nonlocal meanRise, minRise, maxRise, nRising
            meanRise = F32(meanRise + s)
            if nRising == 0:
                minRise = maxRise = s
            else:
                minRise, maxRise = min(minRise, s), max(maxRise, s)
            nRising += 1

        def fall(s):
            nonlocal meanFall, minFall, maxFall, nFalling
            meanFall = F32(meanFall + s)
            if nFalling == 0:
                minFall = maxFall = s
            else:
                minFall, maxFall = min(minFall, s), max(maxFall, s)
            nFalling += 1

        for el in lst:
            if el[0] == 0:
                lastMin, lastMinPos = el[2], el[1]
                if lastMinPos - lastMaxPos > 0:
                    fall(F32(F32(lastMax - lastMin) / F32(F32(lastMinPos - lastMaxPos) * T)))
                    lastIsMax = 0
            else:
                lastMax, lastMaxPos = el[2], el[1]
                if lastMaxPos - lastMinPos > 0:
                    rise(F32(F32(lastMax - lastMin) / F32(F32(lastMaxPos - lastMinPos) * T)))
                    lastIsMax = 1
        if lastIsMax == 1:
            if N - 1 - lastMaxPos > 0:
                fall(F32(F32(x[N - 1] - lastMax) / F32(F32(N - 1 - lastMaxPos) * T)))
        elif lastIsMax == 0:
            if N - 1 - lastMinPos > 0:
                rise(F32(F32(x[N - 1] - lastMin) / F32(F32(N - 1 - lastMinPos) * T)))
        else:
            s_ = F32(F32(x[N - 1] - x[0]) / F32(N))
            if s_ > 0:
                meanRise = maxRise = minRise = s_
                nRising = 1
            elif s_ < 0:
                meanFall = maxFall = minFall = s_
                nFalling = 1
        if nRising > 1:
            meanRise = F32(meanRise / F32(nRising))
        if nFalling > 1:
            meanFall = F32(meanFall / F32(nFalling))
        lastMax = lastMin = F32(x[0])
        lastMaxPos = lastMinPos = 0
        for el in lst:
            if el[0] == 0:
                lastMin, lastMinPos = el[2], el[1]
                if lastMinPos - lastMaxPos > 0:
                    s_ = F32(F32(lastMax - lastMin) / F32(F32(lastMinPos - lastMaxPos) * T))
                    d = F32(s_ - meanFall)
                    sdFall = F32(sdFall + F32(d * d))
            else:
                lastMax, lastMaxPos = el[2], el[1]
                if lastMaxPos - lastMinPos:
                    s_ = F32(F32(lastMax - lastMin) / F32(F32(lastMaxPos - lastMinPos) * T))
                    d = F32(s_ - meanRise)
                    sdRise = F32(sdRise + F32(d * d))
        if nRising > 1:
            sdRise = F32(sdRise / F32(nRising))
        if nFalling > 1:
            sdFall = F32(sdFall / F32(nFalling))
        sdRise, sdFall = sq(sdRise), sq(sdFall)
    nrm = _norm(c["norm"], c["norm_set"], spec.master_norm)
    P_ = F32(period)
    if nrm == SECOND:
        peakDist, peakSdDist, minDist, minSdDist = F32(peakDist * P_), F32(peakSdDist * P_), F32(minDist * P_), F32(minSdDist * P_)
    elif nrm == SEGMENT:
        peakDist, peakSdDist, minDist, minSdDist = F32(peakDist / F32(N)), F32(peakSdDist / F32(N)), F32(minDist / F32(N)), F32(minSdDist / F32(N))
    lim = (lambda v: _ratio_limit(v)) if c["doRatioLimit"] else (lambda v: F32(v))
    limMax = (lambda alt: F32(20.0)) if c["doRatioLimit"] else (lambda alt: F32(alt))
    unity = (lambda v: min(max(F32(v), F32(-1)), F32(1))) if c["doRatioLimit"] else (lambda v: F32(v))
    v = {}
    v["numPeaks"] = F32(F32(nPeaks) / F32(F32(N) * P_)) if nrm == SECOND else F32(nPeaks)
    v["meanPeakDist"], v["meanPeakDistDelta"], v["peakDistStddev"] = peakDist, F32(0), peakSdDist
    v["peakRangeAbs"] = F32(peakMax - peakMin)
    v["peakRangeRel"] = unity(F32(abs(F32(F32(peakMax - peakMin) / rng)))) if rng != 0 else F32(peakMax - peakMin)
    v["peakMeanAbs"], v["peakMeanMeanDist"] = peakMean, F32(peakMean - mean)
    v["peakMeanRel"] = lim(F32(peakMean / mean)) if mean != 0 else limMax(peakMean)
    v["ptpAmpMeanAbs"] = peakDiff
    v["ptpAmpMeanRel"] = unity(F32(peakDiff / rng)) if rng != 0 else peakDiff
    v["ptpAmpStddevAbs"] = peakSdDiff
    v["ptpAmpStddevRel"] = unity(F32(peakSdDiff / rng)) if rng != 0 else peakSdDiff
    v["minRangeAbs"] = F32(minMax - minMin)
    v["minRangeRel"] = unity(F32(abs(F32(F32(minMax - minMin) / rng)))) if rng != 0 else F32(minMax - minMin)
    v["minMeanAbs"], v["minMeanMeanDist"] = minMean, F32(mean - minMean)
    v["minMeanRel"] = lim(F32(minMean / mean)) if mean != 0 else limMax(minMean)
    v["mtmAmpMeanAbs"] = minDiff
    v["mtmAmpMeanRel"] = unity(F32(minDiff / rng)) if rng != 0 else minDiff
    v["mtmAmpStddevAbs"] = minSdDiff
    v["mtmAmpStddevRel"] = unity(F32(minSdDiff / rng)) if rng != 0 else minSdDiff
    v["meanRisingSlope"], v["maxRisingSlope"], v["minRisingSlope"], v["stddevRisingSlope"] = meanRise, maxRise, minRise, sdRise
    v["meanFallingSlope"], v["maxFallingSlope"], v["minFallingSlope"], v["stddevFallingSlope"] = meanFall, maxFall, minFall, sdFall
    v["covFallingSlope"] = lim(F32(sdFall / meanFall)) if meanFall > 0 else F32(0)
    v["covRisingSlope"] = lim(F32(sdRise / meanRise)) if meanRise > 0 else F32(0)
    return [F32(v[k]) for k in PEAKS2_NAMES if c[k]]


def functionals(spec, rows, period):
    """rows [T, K] float32 (one utterance) -> [K * n_values] float32, values of element e contiguous"""
    rows = np.asarray(rows, np.float32)
    return np.array([v for e in range(rows.shape[1]) for v in contour(spec, rows[:, e], period)], np.float32)


IS09 = Spec(["Extremes", "Regression", "Moments"],
            extremes=dict(max=1, min=1, range=1, maxpos=1, minpos=1, amean=1, maxameandist=0, minameandist=0, norm=FRAME, norm_set=True),
            regression=dict(linregc1=1, linregc2=1, linregerrA=0, linregerrQ=1, qregc1=0, qregc2=0, qregc3=0, qregerrA=0, qregerrQ=0, centroid=0,
                            oldBuggyQerr=1, normInputs=0, normRegCoeff=0),
            moments=dict(variance=0, stddev=1, skewness=1, kurtosis=1, amean=0))
"""config/is09-13/IS09_emotion_core.func.conf.inc: the INTERSPEECH 2009 Emotion Challenge functionals (12 per contour)"""
