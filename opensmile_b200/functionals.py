"""ctypes mirror of include/osm_b200_functionals.h: cFunctionals (full-input mode) on the GPU.

The reference computes these per-utterance summaries with cFunctionals and its cFunctional* sub-components
(src/functionals/functionals.cpp:284-330); here one warp per (utterance, LLD element) computes them from LLD rows that are
resident in HBM.  No CPU fallback."""
import ctypes as C

import numpy as np

from . import capi

i32, f64, f32 = C.c_int32, C.c_double, C.c_float
F_EXTREMES, F_MEANS, F_MOMENTS, F_PERCENTILES, F_REGRESSION, F_TIMES, F_LPC, F_SEGMENTS, F_PEAKS2, F_ONSET, F_PEAKS, F_CROSSINGS, F_SAMPLES, F_DCT = range(14)
TYPE_BY_NAME = {"Extremes": F_EXTREMES, "Means": F_MEANS, "Moments": F_MOMENTS, "Percentiles": F_PERCENTILES, "Regression": F_REGRESSION,
                "Times": F_TIMES, "Lpc": F_LPC, "Segments": F_SEGMENTS, "Peaks2": F_PEAKS2, "Onset": F_ONSET, "Peaks": F_PEAKS,
                "Crossings": F_CROSSINGS, "Samples": F_SAMPLES, "DCT": F_DCT}
SEG_RELTH, SEG_NONX, SEG_EQX, SEG_NARELTH = 0, 1, 2, 3
SEG_BY_NAME = {"relTh": SEG_RELTH, "nonX": SEG_NONX, "eqX": SEG_EQX, "NArelTh": SEG_NARELTH}
PEAKS2_NAMES = ["numPeaks", "meanPeakDist", "meanPeakDistDelta", "peakDistStddev", "peakRangeAbs", "peakRangeRel", "peakMeanAbs",
                "peakMeanMeanDist", "peakMeanRel", "ptpAmpMeanAbs", "ptpAmpMeanRel", "ptpAmpStddevAbs", "ptpAmpStddevRel", "minRangeAbs",
                "minRangeRel", "minMeanAbs", "minMeanMeanDist", "minMeanRel", "mtmAmpMeanAbs", "mtmAmpMeanRel", "mtmAmpStddevAbs",
                "mtmAmpStddevRel", "meanRisingSlope", "maxRisingSlope", "minRisingSlope", "stddevRisingSlope", "meanFallingSlope",
                "maxFallingSlope", "minFallingSlope", "stddevFallingSlope", "covFallingSlope", "covRisingSlope"]
TIMENORM_UNSET, TIMENORM_SEGMENT, TIMENORM_SECOND, TIMENORM_FRAME = -1, 0, 1, 2


class _Extremes(C.Structure):
    _fields_ = [(n, i32) for n in ("max", "min", "range", "maxpos", "minpos", "amean", "maxameandist", "minameandist", "norm", "normIsSet")]


class _Means(C.Structure):
    _fields_ = [(n, i32) for n in ("amean", "absmean", "qmean", "nzamean", "nzabsmean", "nzqmean", "nzgmean", "nnz", "flatness", "posamean",
                                    "negamean", "posqmean", "posrqmean", "negqmean", "negrqmean", "rqmean", "nzrqmean", "norm", "normIsSet")]


class _Moments(C.Structure):
    _fields_ = [(n, i32) for n in ("variance", "stddev", "skewness", "kurtosis", "amean", "stddevNorm", "doRatioLimit")]


class _Percentiles(C.Structure):
    _fields_ = [(n, i32) for n in ("quartile1", "quartile2", "quartile3", "iqr12", "iqr23", "iqr13")] + \
               [("n_percentile", i32), ("percentile", f64 * 8), ("n_pctlrange", i32), ("pctlrange", (i32 * 2) * 8), ("interp", i32)]


class _Regression(C.Structure):
    _fields_ = [(n, i32) for n in ("linregc1", "linregc2", "linregerrA", "linregerrQ", "qregc1", "qregc2", "qregc3", "qregerrA", "qregerrQ",
                                    "centroid", "centroidNorm", "centroidUseAbsValues", "centroidRatioLimit", "normRegCoeff", "normInputs",
                                    "oldBuggyQerr", "doRatioLimit")]


class _Times(C.Structure):
    _fields_ = [(n, i32) for n in ("upleveltime25", "downleveltime25", "upleveltime50", "downleveltime50", "upleveltime75", "downleveltime75",
                                    "upleveltime90", "downleveltime90", "risetime", "falltime", "leftctime", "rightctime", "duration",
                                    "buggySecNorm", "norm", "normIsSet")]


class _Lpc(C.Structure):
    _fields_ = [(n, i32) for n in ("lpGain", "lpc", "firstCoeff", "order")]


class _Segments(C.Structure):
    _fields_ = [(n, i32) for n in ("numSegments", "meanSegLen", "maxSegLen", "minSegLen", "segLenStddev", "algorithm", "maxNumSeg", "n_thresholds")] + \
               [("thresholds", f32 * 8), ("X", f32)] + [(n, i32) for n in ("XisRel", "segMinLng", "segMinLngIsSet", "pauseMinLng", "norm", "normIsSet")]


class _Peaks2(C.Structure):
    _fields_ = [("value", i32 * 32), ("relThresh", f32), ("absThresh", f32)] + \
               [(n, i32) for n in ("useAbsThresh", "dynRelThresh", "doRatioLimit", "norm", "normIsSet")]


class _Onset(C.Structure):
    _fields_ = [(n, i32) for n in ("onsetPos", "offsetPos", "numOnsets", "numOffsets", "onsetRate")] + \
               [("thresholdOnset", f32), ("thresholdOffset", f32)] + [(n, i32) for n in ("useAbsVal", "norm", "normIsSet")]


class _Peaks(C.Structure):
    _fields_ = [(n, i32) for n in ("numPeaks", "meanPeakDist", "peakMean", "peakMeanMeanDist", "peakDistStddev", "norm", "normIsSet")]


class _Crossings(C.Structure):
    _fields_ = [(n, i32) for n in ("zcr", "mcr", "amean")]


class _Samples(C.Structure):
    _fields_ = [("n_samplepos", i32), ("samplepos", f64 * 16)]


class _Dct(C.Structure):
    _fields_ = [("firstCoeff", i32), ("lastCoeff", i32)]


class Spec(C.Structure):
    _fields_ = [("n_enabled", i32), ("enabled", i32 * 8), ("nonZeroFuncts", i32), ("masterTimeNorm", i32),
                ("functNameAppend", C.c_char * capi.NAME_LEN), ("extremes", _Extremes), ("means", _Means), ("moments", _Moments),
                ("percentiles", _Percentiles), ("regression", _Regression), ("times", _Times), ("lpc", _Lpc), ("segments", _Segments),
                ("peaks2", _Peaks2), ("onset", _Onset), ("peaks", _Peaks), ("crossings", _Crossings),
                ("samples", _Samples), ("dct", _Dct)]


def _bind(L):
    if getattr(L, "_fn_bound", False):
        return L
    vp = C.c_void_p
    L.osm_b200_functionals_defaults.argtypes = [C.POINTER(Spec)]
    L.osm_b200_functionals_defaults.restype = None
    L.osm_b200_functionals_create.argtypes = [C.POINTER(Spec), i32, C.POINTER(C.c_char_p), f64, i32, C.POINTER(vp)]
    L.osm_b200_functionals_destroy.argtypes = [vp]
    L.osm_b200_functionals_destroy.restype = None
    L.osm_b200_functionals_num_values.argtypes = [vp]
    L.osm_b200_functionals_num_elements.argtypes = [vp]
    L.osm_b200_functionals_element_name.argtypes = [vp, i32]
    L.osm_b200_functionals_element_name.restype = C.c_char_p
    i64p = C.POINTER(C.c_int64)
    L.osm_b200_functionals_run_device.argtypes = [vp, vp, i32, i64p, i64p, i32, vp, vp]
    L.osm_b200_functionals_run_host.argtypes = [vp, vp, i32, i64p, i64p, i32, C.c_int64, vp]
    assert L.osm_b200_functionals_sizeof_spec() == C.sizeof(Spec), "ctypes mirror of osm_b200_functionals_spec is out of date"
    L._fn_bound = True
    return L


def default_spec():
    s = Spec()
    _bind(capi.lib()).osm_b200_functionals_defaults(C.byref(s))
    return s


def spec(enabled, non_zero=0, master_norm=TIMENORM_UNSET, name_append="", **sub):
    """enabled: list of functional names in functionalsEnabled order; sub: extremes=dict(...), means=..., moments=..., percentiles=...,
    regression=... with the reference's field names (percentile / pctlrange as lists)"""
    s = default_spec()
    s.n_enabled = len(enabled)
    for i, n in enumerate(enabled):
        s.enabled[i] = TYPE_BY_NAME[n]
    s.nonZeroFuncts, s.masterTimeNorm, s.functNameAppend = non_zero, master_norm, name_append.encode()
    for part, fields in sub.items():
        blk = getattr(s, part)
        for k, v in fields.items():
            if k == "percentile":
                blk.n_percentile = len(v)
                for i, x in enumerate(v):
                    blk.percentile[i] = x
            elif k == "pctlrange":
                blk.n_pctlrange = len(v)
                for i, (a, b) in enumerate(v):
                    blk.pctlrange[i][0], blk.pctlrange[i][1] = a, b
            elif k == "samplepos":
                blk.n_samplepos = len(v)
                for i, x in enumerate(v):
                    blk.samplepos[i] = x
            elif k == "thresholds":
                blk.n_thresholds = len(v)
                for i, x in enumerate(v):
                    blk.thresholds[i] = x
            elif k == "segmentationAlgorithm":
                blk.algorithm = SEG_BY_NAME[v]
            elif part == "peaks2" and k in PEAKS2_NAMES:
                blk.value[PEAKS2_NAMES.index(k)] = v
            else:
                setattr(blk, k, v)
    return s


class Functionals:
    def __init__(self, spec_, in_names, input_period, device=0):
        self._L = _bind(capi.lib())
        self._h = C.c_void_p()
        arr = (C.c_char_p * len(in_names))(*[n.encode() for n in in_names])
        st = self._L.osm_b200_functionals_create(C.byref(spec_), len(in_names), arr, float(input_period), device, C.byref(self._h))
        if st != capi.OK:
            self._h = C.c_void_p()
            raise RuntimeError(capi.last_error())
        self.n_in = len(in_names)

    def close(self):
        if self._h:
            self._L.osm_b200_functionals_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @property
    def num_values(self):
        return self._L.osm_b200_functionals_num_values(self._h)

    @property
    def num_elements(self):
        return self._L.osm_b200_functionals_num_elements(self._h)

    def element_names(self):
        return [self._L.osm_b200_functionals_element_name(self._h, i).decode() for i in range(self.num_elements)]

    def run_host(self, rows, row_offsets, n_rows):
        """rows [R, stride] float32 (host) -> [n_utt, num_elements]"""
        rows = np.ascontiguousarray(rows, np.float32)
        ro = np.ascontiguousarray(row_offsets, np.int64)
        nr = np.ascontiguousarray(n_rows, np.int64)
        out = np.zeros((len(ro), self.num_elements), np.float32)
        i64p = C.POINTER(C.c_int64)
        st = self._L.osm_b200_functionals_run_host(self._h, rows.ctypes.data, rows.shape[1], ro.ctypes.data_as(i64p), nr.ctypes.data_as(i64p),
                                                   len(ro), rows.shape[0], out.ctypes.data)
        if st != capi.OK:
            raise RuntimeError(capi.last_error())
        return out

    def run_device(self, d_rows, row_stride, row_offsets, n_rows, d_out, stream=None):
        """device pointers (ints or objects with data_ptr()); asynchronous"""
        ro = np.ascontiguousarray(row_offsets, np.int64)
        nr = np.ascontiguousarray(n_rows, np.int64)
        i64p = C.POINTER(C.c_int64)
        ptr = lambda b: C.c_void_p(b.data_ptr() if hasattr(b, "data_ptr") else int(b))
        st = self._L.osm_b200_functionals_run_device(self._h, ptr(d_rows), row_stride, ro.ctypes.data_as(i64p), nr.ctypes.data_as(i64p), len(ro),
                                                     ptr(d_out), C.c_void_p(stream or 0))
        if st != capi.OK:
            raise RuntimeError(capi.last_error())
