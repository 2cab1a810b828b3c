"""CPU restatement (numpy) of the formant chain of the GeMAPS graphs:
cTransformFFT output -> cSpecResample -> cLpc (acf) -> cFormantLpc.

TEST INFRASTRUCTURE ONLY (same rules as oracle/osm_oracle.h): nothing in the product imports it.  Pinned against
the UNMODIFIED reference through level taps (tests/configs/formant_taps.conf, scripts/make_golden_formant.py ->
tests/golden/formant_goldens.npz).  Citations are relative to /root/reference/src.  float32 where the reference
computes in FLOAT_DMEM (sequential accumulation order kept), float64 where it computes in double.  The polynomial
roots come from numpy (LAPACK eigenvalues of the companion matrix) instead of the reference's own QR iteration on
the balanced companion matrix (smileutil/zerosolve.cpp): the same roots to ~1e-12, and cFormantLpc sorts the
resulting formants by frequency, so the order in which a solver returns them does not matter.
"""
import ctypes as C

import numpy as np

from . import oracle

f32 = np.float32


def fft_frames(pcm, fe, n_chan=1):
    """packed real FFT of every frame (dspcore/transformFft.cpp:165-223, packing dspcore/fftsg.c:104-122) -> [T, nfft]"""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    nS = pcm.size // n_chan
    N, H, nfft, T = oracle.geometry(fe, nS)
    L = oracle.lib()
    x = np.zeros(nS, np.float32)
    L.osm_or_pcm16_to_float(pcm.ctypes.data_as(C.POINTER(C.c_int16)), C.c_long(nS), C.c_int(n_chan), oracle._fp(x))
    win = np.zeros(N, np.float64)
    L.osm_or_window_table(C.c_int(fe.win_func), C.c_long(N), C.c_double(fe.win_sigma), C.c_double(fe.win_gain),
                          win.ctypes.data_as(C.POINTER(C.c_double)))
    out = np.zeros((max(T, 0), nfft), np.float32)
    mag = np.zeros(nfft // 2 + 1, np.float32)
    for t in range(max(T, 0)):
        L.osm_or_frame_to_mag(C.byref(fe), oracle._fp(x[t * H:]), C.c_long(N), C.c_long(nfft),
                              win.ctypes.data_as(C.POINTER(C.c_double)), oracle._fp(out[t]), oracle._fp(mag))
    return out


class SpecResample:
    """cSpecResample (dsp/specResample.cpp:97-185) on a zero-padded FFT level; smileDsp_initIrdft / smileDsp_irdft
    (smileutil/smileUtil.c:1752-1820)"""

    def __init__(self, n_in, sample_rate, target_fs, frame_size_sec, last_frame_size_sec):
        bT = 1.0 / sample_rate
        sr = 1.0 / bT
        ratio = target_fs / sr                                                # :112-114
        self.base_period_out = 1.0 / target_fs                                # :120 (set before the adjustment below)
        fs, lfs = frame_size_sec, last_frame_size_sec
        if fs != lfs and lfs != 0.0 and lfs != bT:                            # :150-160 zero-padded input
            n_out0 = np.round(float(n_in) * ratio * lfs / fs)
            new_ratio = n_out0 / (float(n_in) * (lfs / fs))
            if new_ratio != ratio:
                ratio = new_ratio
            nd = float(n_in) * ratio
        else:
            n_out0 = np.round(float(n_in) * ratio)
            new_ratio = n_out0 / float(n_in)
            if new_ratio != ratio:
                ratio = new_ratio
            nd = n_out0
        self.K, self.I, self.nd = int(n_in), int(n_out0), nd
        k_max = min(self.K, self.I)                                           # antiAlias = 1
        if k_max & 1:
            k_max -= 1
        self.k_max = k_max
        i = np.arange(self.I, dtype=np.float64)[:, None]
        k2 = np.arange(1, k_max // 2, dtype=np.float64)[None, :]              # k = 2, 4, .. k_max-2
        kn = 2.0 * np.pi * (k2 * i) / nd
        self.cos = np.cos(kn).astype(f32)
        self.sin = np.sin(kn).astype(f32)
        self.nyq = np.cos((2.0 * np.pi * ((self.K // 2) * i[:, 0])) / nd).astype(f32) if self.I >= self.K else None

    def __call__(self, a):
        """a = one packed FFT frame [K] float32 -> [I] float32 (float accumulation in the reference's order)"""
        out = np.full(self.I, a[0], f32)
        if self.nyq is not None:
            out = (out + a[1] * self.nyq).astype(f32)
        for j in range(self.cos.shape[1]):
            k = 2 * (j + 1)
            out = (out + a[k] * self.cos[:, j]).astype(f32)
            out = (out + a[k + 1] * self.sin[:, j]).astype(f32)
        return (out / f32(self.K // 2)).astype(f32)


def autocorr(x, lags):
    """smileDsp_autoCorr (smileutil/smileUtil.c:1560-1569): float accumulation over i = lag .. n-1"""
    x = np.asarray(x, f32)
    n = x.size
    out = np.zeros(lags, f32)
    for lag in range(lags):
        acc = f32(0)
        prod = (x[lag:] * x[:n - lag]).astype(f32)
        for v in prod:
            acc = f32(acc + v)
        out[lag] = acc
    return out


def lpc_acf(r, p):
    """smileDsp_calcLpcAcf (smileutil/smileUtil.c:1572-1627): Durbin recursion in float -> (a[p], gain)"""
    a = np.zeros(p + 1, f32)
    if r[0] == 0:
        return a[:p], f32(0)
    e = f32(r[0])
    for m in range(1, p + 1):
        s = f32(f32(1.0) * r[m])
        for i in range(1, m):
            s = f32(s + f32(a[i - 1] * r[m - i]))
        k_m = f32(f32(f32(-1.0) / e) * s)
        a[m - 1] = k_m
        for i in range(1, m // 2 + 1):
            x = a[i - 1]
            a[i - 1] = f32(a[i - 1] + f32(k_m * a[m - i - 1]))
            if i < m // 2 or (m & 1) == 1:
                a[m - i - 1] = f32(a[m - i - 1] + f32(k_m * x))
        e = f32(e * f32(f32(1.0) - f32(k_m * k_m)))
        if e == 0:
            a[m:] = 0
            break
    return a[:p], e


def formants_from_lpc(a, T, n_formants, min_f, max_f):
    """cFormantLpc::processVector, root branch (lld/formantLpc.cpp:255-301) + smileMath_complexIntoUnitCircle
    (smileUtil.c:992-1004) + smileDsp_lpcrootsToFormants (:2019-2054) -> (freq[n], bandwidth[n]) float32"""
    p = len(a)
    poly = np.concatenate([-np.asarray(a, np.float64)[::-1], [1.0]])          # ascending powers, leading 1 (:258-262)
    roots = np.roots(poly[::-1])
    out = np.abs(roots) > 1.0
    roots = np.where(out, 1.0 / np.conj(np.where(out, roots, 1.0)), roots)    # 1 / conj(root)
    sp_pi = T * np.pi
    hi = max_f
    if hi < min_f or hi > 1.0 / T:
        hi = 0.5 / T - min_f
    fc, bc = [], []
    for r in roots:
        if r.imag < 0:
            continue
        f = abs(np.arctan2(r.imag, r.real)) / (2.0 * sp_pi)
        if min_f <= f <= hi:
            fc.append(f)
            bc.append(-np.log(abs(r)) / sp_pi)
            if len(fc) >= n_formants:
                break
    order = np.argsort(np.array(fc), kind="stable") if fc else []
    freq = np.zeros(n_formants, np.float64)
    bw = np.zeros(n_formants, np.float64)
    for j, o in enumerate(order):
        freq[j], bw[j] = fc[o], bc[o]
    return freq.astype(f32), bw.astype(f32)


def gemaps_formant_chain(pcm, sample_rate=16000.0, taps=False):
    """config/gemaps/v01b/GeMAPSv01b_core.lld.conf.inc:43-58,250-286 -> [T, 10] = formantFreqLpc[1..5] | formantBandwidthLpc[1..5]"""
    fe = oracle.frontend(sample_rate, 0.020, 0.010, win="ham", zero_pad_symmetric=1)
    spec = fft_frames(pcm, fe)
    N, H, nfft, T = oracle.geometry(fe, len(pcm))
    fs_sec = oracle.lib().osm_or_fft_frame_size_sec(C.byref(fe))
    rs = SpecResample(nfft, sample_rate, 11000.0, fs_sec, 0.020)
    res = np.stack([rs(a) for a in spec]) if T > 0 else np.zeros((0, rs.I), f32)
    lpcs = np.zeros((res.shape[0], 11), f32)
    fmt = np.zeros((res.shape[0], 10), f32)
    for t, x in enumerate(res):
        a, _ = lpc_acf(autocorr(x, 12), 11)
        lpcs[t] = a
        f, b = formants_from_lpc(a, rs.base_period_out, 5, 50.0, 5450.0)
        fmt[t, :5], fmt[t, 5:] = f, b
    return (fmt, res, lpcs) if taps else fmt
