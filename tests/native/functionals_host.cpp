// Host build of the order-dependent functionals (opensmile_b200/csrc/functionals_seq.cuh) for the CPU tests: the same statements
// the device executes on lane 0 of a contour's warp; what the warp does in parallel (counting for cFunctionalTimes is not here --
// it has no order-dependent part --, the ordered compaction of the local extrema, one lane per autocorrelation lag) is a plain
// loop.  Test infrastructure: compared with oracle/functionals_oracle.py and the reference's rows (tests/test_functionals_cpu.py).
#include <vector>

#include "../../opensmile_b200/csrc/functionals_seq.cuh"

using namespace osm;

extern "C" {

// x: filtered contour; mn / mx / mean as the reference hands them to the sub-components; returns the number of values written
int fsh_segments(const osm_b200_functionals_spec *s, const float *x, long N, float mn, float mx, float period, int timeNorm, float *out)
{
  std::vector<float> lens((size_t)s->segments.maxNumSeg + 1);
  return fseq::segments(s->segments, x, N, mn, mx, period, timeNorm, lens.data(), out);
}

int fsh_peaks2(const osm_b200_functionals_spec *s, const float *x, long N, float mn, float mx, float mean, float period, int timeNorm, float *out)
{
  std::vector<float> ly((size_t)N + 1);
  std::vector<int> lx((size_t)N + 1);
  int nl = 0;
  for (long i = 2; i < N - 2; i++) {                                   // functionalPeaks2.cpp:320-327
    if (x[i] > x[i - 1] && x[i] > x[i + 1]) { ly[nl] = x[i]; lx[nl++] = (int)(i << 1) | 1; }
    else if (x[i] < x[i - 1] && x[i] < x[i + 1]) { ly[nl] = x[i]; lx[nl++] = (int)(i << 1); }
  }
  return fseq::peaks2(s->peaks2, x, N, mn, mx, mean, period, timeNorm, ly.data(), lx.data(), nl, out);
}

int fsh_lpc(const osm_b200_functionals_spec *s, const float *x, long N, float *out)
{
  float acf[OSM_B200_F_MAX_LPC + 1];
  for (int lag = 0; lag <= s->lpc.order; lag++) {                      // smileDsp_autoCorr (smileUtil.c:1560-1569)
    float acc = 0.0f;
    for (long i = lag; i < N; i++) acc = acc + x[i] * x[i - lag];
    acf[lag] = acc;
  }
  return fseq::lpc(s->lpc, acf, N, out);
}

int fsh_onset(const osm_b200_functionals_spec *s, const float *x, long N, float period, int timeNorm, float *out)
{
  return fseq::onset(s->onset, x, N, period, timeNorm, out);
}

int fsh_peaks(const osm_b200_functionals_spec *s, const float *x, long N, float period, int timeNorm, float *out)
{
  std::vector<int> dists((size_t)N + 1);
  return fseq::peaks(s->peaks, x, N, period, timeNorm, dists.data(), out);
}

int fsh_crossings(const osm_b200_functionals_spec *s, const float *x, long N, float *out)
{
  return fseq::crossings(s->crossings, x, N, out);
}

}  // extern "C"
