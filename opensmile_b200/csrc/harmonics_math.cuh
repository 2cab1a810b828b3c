// harmonics_math.cuh -- per-frame arithmetic of cHarmonics (lld/harmonics.cpp) for the switch set of the GeMAPS graphs:
// HNR from the autocorrelation at the F0 lag, harmonic peaks of the magnitude spectrum around multiples of F0, their log
// magnitudes relative to the fundamental, harmonic differences (H1-H2, H1-A3 ...) and formant amplitudes.
// Written once for the device (harmonics.cu) and for a host build of the same statements (tests/native/formant_host.cpp).
// Citations relative to /root/reference/src.  Compile with FMA contraction off.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OSM_HM_HD __host__ __device__ __forceinline__
#else
#define OSM_HM_HD inline
#endif

namespace osm {
namespace hm {

constexpr int kMaxHarmonics = 128;
constexpr int kMaxDiffs = 4;
constexpr int kMaxFormants = 8;

struct Harm { int bin; float fi, mag, magi, lr; };           // sF0Harmonic: bin, freqInterpolated, magnitude, magnitudeInterpolated, magnitudeLogRelF0

// isPeak (lld/harmonics.cpp:369-390)
template <class X>
OSM_HM_HD bool is_peak(const X &x, int N, int n)
{
  if (n >= N || n < 0) return false;
  if (n + 1 < N) {
    if (n > 0) return x(n) > x(n - 1) && x(n) > x(n + 1);
    return x(0) > x(1);
  }
  return n > 0 && x(n) > x(n - 1);
}

// freqToBin (:403-415) on the linear axis frq[b] = b * binHz.  The reference walks b = start, start+1, .. until
// frq[b] > freq and then picks the closer of b-1 and b (0 if it runs off the axis).  On a linear axis the first such b is
// found in closed form; the two fix-up loops re-establish the reference's exact predicate (double product, strict >)
// where the division rounds the other way.
OSM_HM_HD int freq_to_bin(double binHz, int nb, float freq, int start)
{
  const double fr = (double)freq;
  long b0 = (long)floor(fr / binHz);                       // last bin with frq[b0] <= freq, up to rounding
  if (b0 < -1) b0 = -1;
  if (b0 > (long)nb) b0 = (long)nb;
  while (b0 >= 0 && (double)b0 * binHz > fr) b0--;
  while (b0 + 1 <= (long)nb && (double)(b0 + 1) * binHz <= fr) b0++;
  long b = b0 + 1;
  if (b < (long)start) b = (long)start;
  if (b >= (long)nb) return 0;
  const double fb = (double)b * binHz;
  return (fb - fr > fr - (double)(b - 1) * binHz) ? (int)b - 1 : (int)b;
}

// smileMath_quadFrom3pts (smileutil/smileUtil.c:1009-1034): vertex (x, y) of the parabola through three points
OSM_HM_HD void quad3(double x1, double y1, double x2, double y2, double x3, double y3, double *xo, double *yo)
{
  const double den = x1 * x1 * x2 + x2 * x2 * x3 + x3 * x3 * x1 - x3 * x3 * x2 - x2 * x2 * x1 - x1 * x1 * x3;
  if (den != 0.0) {
    const double a = (y1 * x2 + y2 * x3 + y3 * x1 - y3 * x2 - y2 * x1 - y1 * x3) / den;
    const double b = (x1 * x1 * y2 + x2 * x2 * y3 + x3 * x3 * y1 - x3 * x3 * y2 - x2 * x2 * y1 - x1 * x1 * y3) / den;
    const double c = (x1 * x1 * x2 * y3 + x2 * x2 * x3 * y1 + x3 * x3 * x1 * y2 - x3 * x3 * x2 * y1 - x2 * x2 * x1 * y3 - x1 * x1 * x3 * y2) / den;
    if (a != 0.0) { const double x = -b / (2.0 * a); *xo = x; *yo = c - a * x * x; return; }
  }
  if (y1 > y2 && y1 > y3) { *xo = x1; *yo = y1; return; }
  if (y2 > y1 && y2 > y3) { *xo = x2; *yo = y2; return; }
  if (y3 > y1 && y3 > y2) { *xo = x3; *yo = y3; return; }
  *xo = x1; *yo = y1;
}

// postProcessHarmonics with logRelMagnitude (:550-588), split in two so that a warp can share the work:
// log magnitude of harmonic i >= 1 relative to the fundamental (log10 of a float argument is the float function in the
// reference's build: log10f, widened afterwards) ...
OSM_HM_HD float log_rel(float magi, bool logRel, float m0log)
{
  if (!logRel) return -201.0f;
  if (magi > 0.0f) {
    const double t = (double)log10f(magi);
    const float v = (float)(20.0 * (t - (double)m0log));
    return v >= -200.0f ? v : -200.0f;
  }
  return -200.0f;
}
// ... and the removal of duplicates, which compares with the (possibly already cleared) predecessor: sequential
OSM_HM_HD void dedup(Harm *H, int nHarm)
{
  for (int i = 1; i < nHarm; i++)
    if (H[i].bin == H[i - 1].bin) H[i] = Harm{0, 0.0f, 0.0f, 0.0f, -201.0f};
}
OSM_HM_HD void post_process(Harm *H, int nHarm)
{
  const float m0 = H[0].mag;
  const bool logRel = m0 != 0.0f;
  const float m0log = logRel ? log10f(m0) : 0.0f;
  H[0].lr = 0.0f;                                             // unconditional in the reference (:561)
  for (int i = 1; i < nHarm; i++) H[i].lr = log_rel(H[i].magi, logRel, m0log);
  dedup(H, nHarm);
}

// one harmonic of findHarmonicPeaks, branch with a frequency axis (:476-545): `last` = the previous harmonic's candidate
// bin (lower bound of this one's search), returns this harmonic's candidate bin.  M(b) = magnitude of bin b.
template <class M>
OSM_HM_HD int find_one_harmonic(float pitch, const M &mag, int nb, double binHz, int i, int last, int first, Harm *out)
{
    Harm h{-1, 0.0f, 0.0f, 0.0f, -201.0f};
    const int cand = freq_to_bin(binHz, nb, (float)(i + 1) * pitch, last);
    if (cand >= nb) { *out = h; return last; }
    int peak = -1;
    if (is_peak(mag, nb, cand)) peak = cand;
    else {
      int cl = cand - 1, cr = cand + 1;
      const int lo = freq_to_bin(binHz, nb, ((float)i + 0.5f) * pitch, last);
      const int hi = freq_to_bin(binHz, nb, ((float)i + 1.5f) * pitch, cand);
      while ((cl >= lo || cr <= hi) && peak == -1) {
        if (cr <= hi) { if (is_peak(mag, nb, cr)) { peak = cr; break; } cr++; }
        if (cl >= lo) { if (is_peak(mag, nb, cl)) { peak = cl; break; } cl--; }
      }
    }
    if (peak >= first && peak < nb - 1) {
      h.bin = peak;
      h.mag = mag(peak);
      double x, y;
      quad3((double)(peak - 1) * binHz, (double)mag(peak - 1), (double)peak * binHz, (double)mag(peak), (double)(peak + 1) * binHz,
            (double)mag(peak + 1), &x, &y);
      h.fi = (float)x; h.magi = (float)y;
    } else h.bin = cand;
    *out = h;
    return cand;
}

// findHarmonicPeaks + postProcessHarmonics, sequential form (host build, and the kernel's fallback)
template <class M>
OSM_HM_HD void find_harmonics(float pitch, const M &mag, int nb, double binHz, int nHarm, Harm *H)
{
  int last = freq_to_bin(binHz, nb, 0.5f * pitch, 1);
  const int first = freq_to_bin(binHz, nb, 0.5f * pitch, last);
  for (int i = 0; i < nHarm; i++) last = find_one_harmonic(pitch, mag, nb, binHz, i, last, first, &H[i]);
  post_process(H, nHarm);
}

// the strongest harmonic within +-20 % of a formant frequency (getFormantAmplitudeIndices, :714-741); -1 = none
OSM_HM_HD int formant_harmonic(const Harm *H, int nHarm, float f)
{
  const float lo = 0.8f * f, hi = 1.2f * f;
  int best = -1;
  float bm = 0.0f;
  for (int h = 0; h < nHarm; h++)
    if (lo <= H[h].fi && H[h].fi <= hi && H[h].mag > bm) { best = h; bm = H[h].mag; }
  return best;
}

struct Diff { int h1formant, h1idx, h2formant, h2idx; };     // "H1-A3" = {-1, 1, 3, -1} (:84-160)

// one harmonic difference in log scale (:840-880): fa[k-1] = harmonic index of formant k
OSM_HM_HD float harmonic_difference(const Harm *H, int nHarm, const int *fa, int nFa, Diff d)
{
  int i1 = d.h1idx, i2 = d.h2idx;
  if (d.h1formant > 0) i1 = d.h1formant <= nFa ? fa[d.h1formant - 1] : -1;
  if (d.h2formant > 0) i2 = d.h2formant <= nFa ? fa[d.h2formant - 1] : -1;
  const bool ok1 = i1 >= 0 && i1 < nHarm, ok2 = i2 >= 0 && i2 < nHarm;
  float v;
  if (ok1 && ok2) v = H[i1].lr - H[i2].lr;
  else if (ok1) v = H[i1].lr - 201.0f;
  else if (ok2) v = (float)(-201.0 - (double)H[i2].lr);
  else return 0.0f;
  return v < -201.0f ? -201.0f : (v > 201.0f ? 201.0f : v);
}

// getClosestPeak (:632-665) on a lazily evaluated sequence A(j), j = 0 .. N-1
template <class A>
OSM_HM_HD int closest_peak(const A &x, int N, int idx)
{
  if (idx >= N) {
    // a lag beyond the sequence (F0 below fs / N, never produced by the pitch chain's range): the reference's outward walk
    // then only ever finds peaks below N, highest first, and ends in an out-of-bounds read when there is none; bounded here
    for (int j = N - 1; j > 0; j--) if (is_peak(x, N, j)) return j;
    return 0;
  }
  if (is_peak(x, N, idx)) return idx;
  int o = 1;
  while (idx - o > 0 || idx + o < N - 1) {
    if (idx - o > 0 && is_peak(x, N, idx - o)) return idx - o;
    if (idx + o < N - 1 && is_peak(x, N, idx + o)) return idx + o;
    o++;
  }
  const float x0 = x(0), xi = x(idx), xl = x(N - 1);
  if (x0 > xi && xl <= xi) return 0;
  if (x0 <= xi && xl > xi) return N - 1;
  if (x0 > xi && xl > xi) return idx < N / 2 ? 0 : N - 1;
  return idx;
}

// computeAcfHnr_dB (:690-712) from acf[0] and acf[ref]
OSM_HM_HD float hnr_db(float a0, float ar)
{
  double hnr = (double)a0 - (double)ar;
  hnr = hnr == 0.0 ? 10e10 : (double)ar / hnr;
  if (hnr > 10e10) return (float)(10.0 * log10(10e10));
  if (hnr < 10e-10) return (float)(10.0 * log10(10e-10));
  return (float)(10.0 * log10(hnr));
}

}  // namespace hm
}  // namespace osm
