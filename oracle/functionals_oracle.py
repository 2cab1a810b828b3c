"""CPU restatement of cFunctionals (frameMode = full) with the sub-components cFunctionalExtremes, cFunctionalMeans,
cFunctionalMoments, cFunctionalPercentiles and cFunctionalRegression -- numpy, float64 accumulators like the reference.

TEST INFRASTRUCTURE ONLY (tests/, smoke, bench parity).  Pinned against the unmodified reference's -csvoutput / -arffoutput
rows: tests/golden/functionals_goldens.npz (scripts/make_golden_functionals.py), tests/test_functionals_cpu.py.

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
                 percentiles=None, regression=None):
        self.enabled, self.non_zero, self.master_norm, self.name_append = list(enabled), non_zero, master_norm, name_append
        self.extremes = dict(max=1, min=1, range=1, maxpos=1, minpos=1, amean=0, maxameandist=1, minameandist=1, norm=FRAME, norm_set=False)
        self.extremes.update(extremes or {})
        self.means = dict(amean=1, absmean=1, qmean=1, nzamean=1, nzabsmean=1, nzqmean=1, nzgmean=1, nnz=1, flatness=0, posamean=0,
                          negamean=0, posqmean=0, posrqmean=0, negqmean=0, negrqmean=0, rqmean=0, nzrqmean=0, norm=FRAME, norm_set=False)
        self.means.update(means or {})
        self.moments = dict(variance=1, stddev=1, skewness=1, kurtosis=1, amean=0, stddevNorm=0)
        self.moments.update(moments or {})
        self.percentiles = dict(quartile1=0, quartile2=0, quartile3=0, iqr12=0, iqr23=0, iqr13=0, percentile=[], pctlrange=[], interp=1)
        self.percentiles.update(percentiles or {})
        self.regression = dict(linregc1=1, linregc2=1, linregerrA=1, linregerrQ=1, qregc1=1, qregc2=1, qregc3=1, qregerrA=1, qregerrQ=1,
                               centroid=1, centroidNorm=SEGMENT, centroidUseAbsValues=1, normRegCoeff=0, normInputs=0, oldBuggyQerr=1)
        self.regression.update(regression or {})


EXT_NAMES = ["max", "min", "range", "maxPos", "minPos", "amean", "maxameandist", "minameandist"]
EXT_KEYS = ["max", "min", "range", "maxpos", "minpos", "amean", "maxameandist", "minameandist"]
MEAN_NAMES = ["amean", "absmean", "qmean", "nzamean", "nzabsmean", "nzqmean", "nzgmean", "nnz", "flatness", "posamean", "negamean",
              "posqmean", "posrqmean", "negqmean", "negrqmean", "rqmean", "nzrqmean"]
MOM_NAMES = ["variance", "stddev", "skewness", "kurtosis", "amean"]
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
            ii = np.arange(N, dtype=np.float64)
            num, num2 = (xd * ii).sum(), (xd * ii * ii).sum()
            asum = float(mean) * Nd
            if r["centroidUseAbsValues"]:
                asa = np.abs(xd).sum()
                centroid = (np.abs(xd) * ii).sum() / asa if asa != 0 else 0.0
            else:
                centroid = num / asum if asum != 0 else 0.0
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
    assert len(out) == nvals
    return out


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
