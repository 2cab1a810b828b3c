// functionals_seq.cuh -- the order-dependent functionals of cFunctionals, written once for the device (functionals.cu: lane 0 of
// the contour's warp) and for a host build of the same statements (tests/native/functionals_host.cpp, compared on the CPU with
// oracle/functionals_oracle.py and the reference's rows):
//   cFunctionalSegments  relTh / nonX / eqX   src/functionals/functionalSegments.cpp:240-262, :305-367, :658-797, :800-960
//   cFunctionalPeaks2    pruning passes + statistics  src/functionals/functionalPeaks2.cpp:330-915
//   cFunctionalLpc       Durbin on the contour's autocorrelation  src/functionals/functionalLpc.cpp:98-125
// All arithmetic is FLOAT_DMEM = float, statement by statement (compile with FMA contraction off: -fmad=false /
// -ffp-contract=off).  `x` is the filtered contour (nonZeroFuncts applied), N its length.
#pragma once
#include <math.h>

#include "../../include/osm_b200_functionals.h"

#ifdef __CUDACC__
#define OSM_FS_HD __host__ __device__ __forceinline__
#else
#define OSM_FS_HD inline
#endif

namespace osm {
namespace fseq {

// smileMath_ratioLimit (smileutil/smileUtil.c:602-614) with smileMath_tanh / smileMath_logistic (:590-600)
OSM_FS_HD float tanh_dmem(float a)
{
  const float z = 2.0f * a;
  const float lim = (float)log((double)3.402823466e+38f);
  float lg;
  if (z > lim) lg = 1.0f;
  else if (z < -lim) lg = 0.0f;
  else lg = (float)(1.0 / (1.0 + exp(-(double)z)));
  return 2.0f * lg - 1.0f;
}
OSM_FS_HD float ratio_limit(float x, float limit1, float excess)
{
  if (x > limit1) return tanh_dmem((float)((sqrt((double)x - (double)limit1 + 1.0) - 1.0) / ((double)excess * 0.5))) * excess + limit1;
  if (x < -limit1) return tanh_dmem((float)(-(sqrt(-1.0 * ((double)x + (double)limit1) + 1.0) - 1.0) / ((double)excess * 0.5))) * excess - limit1;
  return x;
}

// ------------------------------------------------------------------------------------------------------------------------
// cFunctionalSegments
// ------------------------------------------------------------------------------------------------------------------------
struct SegState {
  long nSeg, sumLen, maxLen, minLen;
  int maxNumSeg;
  float *lens;          // segment lengths (as float: exact for lengths < 2^24), capacity maxNumSeg
};
OSM_FS_HD long seg_add(SegState &s, long i, long last)            // addNewSegment :240-262
{
  const long L = i - last;
  if (s.nSeg < s.maxNumSeg) {
    s.sumLen += L;
    s.lens[s.nSeg++] = (float)L;
    if (L > s.maxLen) s.maxLen = L;
    if (s.minLen == 0 || L < s.minLen) s.minLen = L;
  }
  return i;
}

// writes the enabled values to out, returns their number.  lens: scratch of maxNumSeg floats
template <class Spec>
OSM_FS_HD int segments(const Spec &g, const float *x, long N, float mn, float mx, float period, int timeNorm, float *lens, float *out)
{
  SegState st;
  st.nSeg = st.sumLen = st.maxLen = st.minLen = 0; st.maxNumSeg = g.maxNumSeg; st.lens = lens;
  const float range = mx - mn;
  if (g.algorithm == OSM_B200_SEG_RELTH) {
    float th[OSM_B200_F_MAX_THRESH];
    for (int j = 0; j < g.n_thresholds; j++) th[j] = mn + range * g.thresholds[j];
    long segMinLng = g.segMinLng < 1 ? 1 : g.segMinLng;
    if (!g.segMinLngIsSet) { segMinLng = N / g.maxNumSeg - 1; if (segMinLng < 2) segMinLng = 2; }
    const long ravgLng = 3;
    long lastSeg = -segMinLng / 2;
    float ravg = 0.0f, raLast = 0.0f;
    for (long i = 0; i < N; i++) {
      ravg = ravg + x[i];
      if (i >= ravgLng) ravg = ravg - x[i - ravgLng];
      const float ra = ravg / (float)((i + 1 < ravgLng) ? i + 1 : ravgLng);
      bool cross = false;
      for (int j = 0; j < g.n_thresholds; j++)
        if ((ra > th[j] && raLast <= th[j]) || (ra < th[j] && raLast >= th[j])) cross = true;
      raLast = ra;
      if (cross && i - lastSeg > segMinLng) lastSeg = seg_add(st, i, lastSeg);
    }
  } else if (g.algorithm == OSM_B200_SEG_NARELTH) {                  // process_SegThreshNoavg :369-413
    float th[OSM_B200_F_MAX_THRESH];
    for (int j = 0; j < g.n_thresholds; j++) th[j] = mn + range * g.thresholds[j];
    long segMinLng = g.segMinLng < 1 ? 1 : g.segMinLng;
    if (!g.segMinLngIsSet) { segMinLng = N / g.maxNumSeg - 1; if (segMinLng < 2) segMinLng = 2; }
    long lastSeg = -segMinLng / 2;
    for (long i = 1; i < N; i++) {
      bool cross = false;
      for (int j = 0; j < g.n_thresholds; j++)
        if ((x[i] > th[j] && x[i - 1] <= th[j]) || (x[i] < th[j] && x[i - 1] >= th[j])) cross = true;
      if (cross && i - lastSeg > segMinLng) lastSeg = seg_add(st, i, lastSeg);
    }
  } else {
    const float X = g.XisRel ? mn + range * g.X : g.X;
    const int segMinLng = g.segMinLng < 1 ? 1 : g.segMinLng, pauseMinLng = g.pauseMinLng < 1 ? 1 : g.pauseMinLng;
    int inSeg = 0, segStart = 0, segEnd = 0;
    long startIdx = 0;
    for (long i = 0; i < N; i++) {
      const bool hit = (g.algorithm == OSM_B200_SEG_NONX) ? (x[i] != X) : (x[i] == X);
      if (hit) {
        if (inSeg == 1) { segEnd = 0; segStart++; if (segStart >= segMinLng) { segStart = 0; inSeg = 2; } }
        else if (inSeg == 0) { segStart++; startIdx = i; inSeg = 1; }
        else segEnd = 0;
      } else {
        if (inSeg == 2) { segStart = 0; segEnd++; if (segEnd >= pauseMinLng) { inSeg = 0; seg_add(st, i - segEnd, startIdx); segEnd = 0; } }
        else if (inSeg == 1) { segEnd++; if (segEnd >= pauseMinLng) { inSeg = 0; segEnd = 0; segStart = 0; } }
      }
    }
    if (inSeg == 2) { segEnd++; seg_add(st, N - segEnd, startIdx); }
  }
  // statistics + output :880-955
  float mean = (st.nSeg > 1) ? (float)st.sumLen / (float)st.nSeg : (float)st.sumLen;
  float dev = 0.0f;
  for (long i = 0; i < st.nSeg; i++) dev = dev + (lens[i] - mean) * (lens[i] - mean);
  if (st.nSeg > 1) { dev = dev / (float)st.nSeg; dev = (float)sqrt((double)dev); } else dev = 0.0f;
  int n = 0;
  const float Nf = (float)N;
  float Tn = 1.0f;
  if (period != 0.0f) Tn = period;
  if (g.numSegments) {
    if (timeNorm == OSM_B200_TIMENORM_SECOND) out[n++] = (float)st.nSeg / (Tn * Nf);
    else if (timeNorm == OSM_B200_TIMENORM_SEGMENT) out[n++] = (float)st.nSeg / (float)g.maxNumSeg;
    else out[n++] = (float)st.nSeg;
  }
  if (timeNorm == OSM_B200_TIMENORM_SEGMENT) {
    if (g.meanSegLen) out[n++] = mean / Nf;
    if (g.maxSegLen) out[n++] = (float)st.maxLen / Nf;
    if (g.minSegLen) out[n++] = (float)st.minLen / Nf;
    if (g.segLenStddev) out[n++] = dev / Nf;
  } else if (timeNorm == OSM_B200_TIMENORM_FRAME) {
    if (g.meanSegLen) out[n++] = mean;
    if (g.maxSegLen) out[n++] = (float)st.maxLen;
    if (g.minSegLen) out[n++] = (float)st.minLen;
    if (g.segLenStddev) out[n++] = dev;
  } else {
    if (g.meanSegLen) out[n++] = mean * Tn;
    if (g.maxSegLen) out[n++] = (float)st.maxLen * Tn;
    if (g.minSegLen) out[n++] = (float)st.minLen * Tn;
    if (g.segLenStddev) out[n++] = dev * Tn;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------------------------------
// cFunctionalPeaks2.  The list of local extrema (step 1, :320-327: built by the caller) is two parallel arrays: ly[k] = value,
// lx[k] = (position << 1) | type (1 = maximum); a removed element gets lx[k] = -1.  The reference's doubly linked list only ever
// unlinks the current element or the remembered last maximum / minimum, so "skip the dead" iteration visits the same elements.
// ------------------------------------------------------------------------------------------------------------------------
template <class Spec>
OSM_FS_HD bool p2_below(const Spec &c, float absT, float diff, float base)            // isBelowThresh :270-294
{
  if (c.dynRelThresh && !c.useAbsThresh) {
    if (base == 0.0f) return diff != 0.0f;
    return fabs((double)(diff / base)) < (double)c.relThresh;
  }
  return diff < absT;
}

template <class Spec>
OSM_FS_HD int peaks2(const Spec &c, const float *x, long N, float mn, float mx, float mean, float period, int timeNorm,
                     float *ly, int *lx, int nl, float *out)
{
  const float range = mx - mn;
  const float absT = c.useAbsThresh ? c.absThresh : c.relThresh * range;
  // ---- step 2a :330-392 ----
  {
    float lastVal = x[0], lastMin = x[0], lastMax = x[0];
    int minFlag = 0, lastMaxPtr = -1;
    for (int k = 0; k < nl; k++) {
      const float y = ly[k];
      if (lx[k] & 1) {
        if (p2_below(c, absT, (float)fabs((double)(y - lastVal)), y < lastVal ? y : lastVal)) {
          if (p2_below(c, absT, y - lastMin, lastMin)) {
            lx[k] = -1;
          } else {
            if ((double)y > (double)lastMax * 1.05) {
              if (lastMaxPtr >= 0) lx[lastMaxPtr] = -1;
              lastMax = y; lastMaxPtr = k;
            } else {
              if (minFlag) { lastMax = y; lastMaxPtr = k; }
              else lx[k] = -1;
            }
            minFlag = 0;
          }
        } else {
          minFlag = 0; lastMax = y; lastMaxPtr = k;
        }
      } else {
        if (!p2_below(c, absT, (float)fabs((double)(y - lastVal)), y < lastVal ? y : lastVal)) { minFlag = 1; lastMin = y; }
      }
      lastVal = y;
    }
  }
  // ---- step 2b :395-412 ----
  {
    float lastMax = x[0];
    for (int k = 0; k < nl; k++) {
      if (lx[k] < 0) continue;
      if (lx[k] & 1) lastMax = ly[k];
      else if (p2_below(c, absT, lastMax - ly[k], ly[k])) lx[k] = -1;
    }
  }
  // ---- step 3 :415-466 ----
  {
    float lastMax = x[0], lastMin = x[0];
    int minFlag = 0, init = 1, lastMinPtr = -1, lastMaxPtr = -1;
    for (int k = 0; k < nl; k++) {
      if (lx[k] < 0) continue;
      const float y = ly[k];
      if (!(lx[k] & 1)) {
        if (!minFlag || init) { lastMin = y; lastMinPtr = k; minFlag = 1; init = 0; }
        else if (y >= lastMin) lx[k] = -1;
        else if (lastMinPtr != k) { lx[lastMinPtr] = -1; lastMinPtr = k; lastMin = y; }
      } else {
        if (minFlag || init) { lastMax = y; lastMaxPtr = k; minFlag = 0; init = 0; }
        else if (y <= lastMax) lx[k] = -1;
        else if (lastMaxPtr != k) { lx[lastMaxPtr] = -1; lastMaxPtr = k; lastMax = y; }
      }
    }
  }
  // ---- statistics, first pass :470-561 ----
  float peakMax = 0.f, peakMin = 0.f, peakDist = 0.f, peakDiff = 0.f, peakMean = 0.f;
  float minMax = 0.f, minMin = 0.f, minDist = 0.f, minDiff = 0.f, minMean = 0.f;
  long nPeakDist = 0, nPeaks = 0, nMinDist = 0, nMins = 0;
  {
    int lastMaxPtr = -1, lastMinPtr = -1;
    for (int k = 0; k < nl; k++) {
      if (lx[k] < 0) continue;
      const float y = ly[k];
      const long pos = lx[k] >> 1;
      if (!(lx[k] & 1)) {
        if (lastMinPtr < 0) { lastMinPtr = k; minMin = y; minMax = y; }
        else {
          nMinDist++;
          minDist = minDist + (float)(pos - (lx[lastMinPtr] >> 1));
          minDiff = minDiff + (float)fabs((double)(y - ly[lastMinPtr]));
          if (minMin > y) minMin = y;
          if (minMax < y) minMax = y;
          lastMinPtr = k;
        }
        minMean = minMean + y; nMins++;
      } else {
        if (lastMaxPtr < 0) { lastMaxPtr = k; peakMin = y; peakMax = y; }
        else {
          nPeakDist++;
          peakDist = peakDist + (float)(pos - (lx[lastMaxPtr] >> 1));
          peakDiff = peakDiff + (float)fabs((double)(y - ly[lastMaxPtr]));
          if (peakMin > y) peakMin = y;
          if (peakMax < y) peakMax = y;
          lastMaxPtr = k;
        }
        peakMean = peakMean + y; nPeaks++;
      }
    }
  }
  if (nPeaks > 1) {
    peakMean = peakMean / (float)nPeaks;
    if (nPeakDist > 1) { peakDist = peakDist / (float)nPeakDist; peakDiff = peakDiff / (float)nPeakDist; }
  }
  if (nMins > 0) {
    minMean = minMean / (float)nMins;
    if (nMinDist > 1) { minDist = minDist / (float)nMinDist; minDiff = minDiff / (float)nMinDist; }
  }
  // ---- second pass :564-610 (the reference takes the peak deviations against the last MINIMUM) ----
  float peakSdDist = 0.f, peakSdDiff = 0.f, minSdDist = 0.f, minSdDiff = 0.f;
  {
    int lastMaxPtr = -1, lastMinPtr = -1;
    for (int k = 0; k < nl; k++) {
      if (lx[k] < 0) continue;
      const float y = ly[k];
      const long pos = lx[k] >> 1;
      if (!(lx[k] & 1)) {
        if (lastMinPtr < 0) lastMinPtr = k;
        else {
          const float a = (float)(pos - (lx[lastMinPtr] >> 1)) - minDist;
          minSdDist = minSdDist + a * a;
          const float b = (float)fabs((double)(y - ly[lastMinPtr])) - minDiff;
          minSdDiff = minSdDiff + b * b;
          lastMinPtr = k;
        }
      } else {
        if (lastMaxPtr < 0) lastMaxPtr = k;
        else {
          // lastMinPtr is set here: after step 3 a minimum lies between any two maxima
          const float a = (float)(pos - (lx[lastMinPtr] >> 1)) - peakDist;
          peakSdDist = peakSdDist + a * a;
          const float b = (float)fabs((double)(y - ly[lastMinPtr])) - peakDiff;
          peakSdDiff = peakSdDiff + b * b;
          lastMaxPtr = k;
        }
      }
    }
  }
  if (nPeakDist > 1) { peakSdDist = peakSdDist / (float)nPeakDist; peakSdDiff = peakSdDiff / (float)nPeakDist; }
  peakSdDist = peakSdDist > 0.0f ? (float)sqrt((double)peakSdDist) : 0.0f;
  peakSdDiff = peakSdDiff > 0.0f ? (float)sqrt((double)peakSdDiff) : 0.0f;
  if (nMinDist > 1) { minSdDist = minSdDist / (float)nMinDist; minSdDiff = minSdDiff / (float)nMinDist; }
  minSdDist = minSdDist > 0.0f ? (float)sqrt((double)minSdDist) : 0.0f;
  minSdDiff = minSdDiff > 0.0f ? (float)sqrt((double)minSdDiff) : 0.0f;
  // ---- slopes :612-745 ----
  float meanRise = 0.f, meanFall = 0.f, minRise = 0.f, maxRise = 0.f, minFall = 0.f, maxFall = 0.f, sdRise = 0.f, sdFall = 0.f;
  int nRising = 0, nFalling = 0;
  bool enabSlope = false;
  for (int k = 22; k < OSM_B200_F_PEAKS2_VALUES; k++) enabSlope = enabSlope || c.value[k];
  if (enabSlope) {
    const float T = period;
    int lastIsMax = -1;
    float lastMax = x[0], lastMin = x[0];
    long lastMaxPos = 0, lastMinPos = 0;
    auto rise = [&](float s) {
      meanRise = meanRise + s;
      if (nRising == 0) { minRise = s; maxRise = s; } else { if (s < minRise) minRise = s; if (s > maxRise) maxRise = s; }
      nRising++;
    };
    auto fall = [&](float s) {
      meanFall = meanFall + s;
      if (nFalling == 0) { minFall = s; maxFall = s; } else { if (s < minFall) minFall = s; if (s > maxFall) maxFall = s; }
      nFalling++;
    };
    for (int k = 0; k < nl; k++) {
      if (lx[k] < 0) continue;
      if (!(lx[k] & 1)) {
        lastMin = ly[k]; lastMinPos = lx[k] >> 1;
        if (lastMinPos - lastMaxPos > 0) { fall((lastMax - lastMin) / ((float)(lastMinPos - lastMaxPos) * T)); lastIsMax = 0; }
      } else {
        lastMax = ly[k]; lastMaxPos = lx[k] >> 1;
        if (lastMaxPos - lastMinPos > 0) { rise((lastMax - lastMin) / ((float)(lastMaxPos - lastMinPos) * T)); lastIsMax = 1; }
      }
    }
    if (lastIsMax == 1) {
      if (N - 1 - lastMaxPos > 0) fall((x[N - 1] - lastMax) / ((float)(N - 1 - lastMaxPos) * T));
    } else if (lastIsMax == 0) {
      if (N - 1 - lastMinPos > 0) rise((x[N - 1] - lastMin) / ((float)(N - 1 - lastMinPos) * T));
    } else {
      const float s = (x[N - 1] - x[0]) / (float)N;
      if (s > 0.0f) { meanRise = maxRise = minRise = s; nRising = 1; }
      else if (s < 0.0f) { meanFall = maxFall = minFall = s; nFalling = 1; }
    }
    if (nRising > 1) meanRise = meanRise / (float)nRising;
    if (nFalling > 1) meanFall = meanFall / (float)nFalling;
    lastMax = x[0]; lastMaxPos = 0; lastMin = x[0]; lastMinPos = 0;
    for (int k = 0; k < nl; k++) {
      if (lx[k] < 0) continue;
      if (!(lx[k] & 1)) {
        lastMin = ly[k]; lastMinPos = lx[k] >> 1;
        if (lastMinPos - lastMaxPos > 0) {
          const float s = (lastMax - lastMin) / ((float)(lastMinPos - lastMaxPos) * T);
          sdFall = sdFall + (s - meanFall) * (s - meanFall);
        }
      } else {
        lastMax = ly[k]; lastMaxPos = lx[k] >> 1;
        if (lastMaxPos - lastMinPos) {
          const float s = (lastMax - lastMin) / ((float)(lastMaxPos - lastMinPos) * T);
          sdRise = sdRise + (s - meanRise) * (s - meanRise);
        }
      }
    }
    if (nRising > 1) sdRise = sdRise / (float)nRising;
    if (nFalling > 1) sdFall = sdFall / (float)nFalling;
    sdRise = sdRise > 0.0f ? (float)sqrt((double)sdRise) : 0.0f;
    sdFall = sdFall > 0.0f ? (float)sqrt((double)sdFall) : 0.0f;
  }
  // ---- normalisation + output :748-905 ----
  const float Nf = (float)N;
  if (timeNorm == OSM_B200_TIMENORM_SECOND) { peakDist = peakDist * period; peakSdDist = peakSdDist * period; minDist = minDist * period; minSdDist = minSdDist * period; }
  else if (timeNorm == OSM_B200_TIMENORM_SEGMENT) { peakDist = peakDist / Nf; peakSdDist = peakSdDist / Nf; minDist = minDist / Nf; minSdDist = minSdDist / Nf; }
  auto lim = [&](float v) { return c.doRatioLimit ? ratio_limit(v, 10.0f, 10.0f) : v; };
  auto limMax = [&](float alt) { return c.doRatioLimit ? 20.0f : alt; };
  auto unity = [&](float v) { if (c.doRatioLimit) { if (v > 1.0f) return 1.0f; if (v < -1.0f) return -1.0f; } return v; };
  float v[OSM_B200_F_PEAKS2_VALUES];
  v[0] = (timeNorm == OSM_B200_TIMENORM_SECOND) ? (float)nPeaks / (Nf * period) : (float)nPeaks;
  v[1] = peakDist; v[2] = 0.0f; v[3] = peakSdDist;
  v[4] = peakMax - peakMin;
  v[5] = range != 0.0f ? unity((float)fabs((double)((peakMax - peakMin) / range))) : peakMax - peakMin;
  v[6] = peakMean; v[7] = peakMean - mean;
  v[8] = mean != 0.0f ? lim(peakMean / mean) : limMax(peakMean);
  v[9] = peakDiff;
  v[10] = range != 0.0f ? unity(peakDiff / range) : peakDiff;
  v[11] = peakSdDiff;
  v[12] = range != 0.0f ? unity(peakSdDiff / range) : peakSdDiff;
  v[13] = minMax - minMin;
  v[14] = range != 0.0f ? unity((float)fabs((double)((minMax - minMin) / range))) : minMax - minMin;
  v[15] = minMean; v[16] = mean - minMean;
  v[17] = mean != 0.0f ? lim(minMean / mean) : limMax(minMean);
  v[18] = minDiff;
  v[19] = range != 0.0f ? unity(minDiff / range) : minDiff;
  v[20] = minSdDiff;
  v[21] = range != 0.0f ? unity(minSdDiff / range) : minSdDiff;
  v[22] = meanRise; v[23] = maxRise; v[24] = minRise; v[25] = sdRise;
  v[26] = meanFall; v[27] = maxFall; v[28] = minFall; v[29] = sdFall;
  v[30] = meanFall > 0.0f ? lim(sdFall / meanFall) : 0.0f;
  v[31] = meanRise > 0.0f ? lim(sdRise / meanRise) : 0.0f;
  int n = 0;
  for (int k = 0; k < OSM_B200_F_PEAKS2_VALUES; k++) if (c.value[k]) out[n++] = v[k];
  return n;
}

// ------------------------------------------------------------------------------------------------------------------------
// cFunctionalLpc: smileDsp_calcLpcAcf (smileutil/smileUtil.c:1572-1627) on acf[0..order]; writes the enabled values
// ------------------------------------------------------------------------------------------------------------------------
template <class Spec>
OSM_FS_HD int lpc(const Spec &l, const float *acf, long N, float *out)
{
  float a[OSM_B200_F_MAX_LPC + 1];
  const int p = l.order;
  for (int i = 0; i <= p; i++) a[i] = 0.0f;
  float gain = 0.0f;
  if (acf[0] != 0.0f) {
    float e = acf[0];
    for (int m = 1; m <= p; m++) {
      float s = 1.0f * acf[m];
      for (int i = 1; i < m; i++) s = s + a[i - 1] * acf[m - i];
      const float km = (-1.0f / e) * s;
      a[m - 1] = km;
      for (int i = 1; i <= m / 2; i++) {
        const float t = a[i - 1];
        a[i - 1] = a[i - 1] + km * a[m - i - 1];
        if (i < m / 2 || (m & 1) == 1) a[m - i - 1] = a[m - i - 1] + km * t;
      }
      e = e * (1.0f - km * km);
      if (e == 0.0f) { for (int i = m; i <= p; i++) a[i] = 0.0f; break; }
    }
    gain = e;
  }
  int n = 0;
  if (l.lpGain) out[n++] = gain / (float)N;
  if (l.lpc) for (int i = l.firstCoeff; i < p; i++) out[n++] = a[i];
  return n;
}

// cFunctionalOnset::process (functionalOnset.cpp:95-153)
template <class Spec>
OSM_FS_HD int onset(const Spec &c, const float *x, long N, float period, int timeNorm, float *out)
{
  long onsetPos = -1, offsetPos = -1, nOnsets = 0, nOffsets = 0;
  int oo = x[0] > c.thresholdOnset ? 1 : 0;
  for (long i = 1; i < N; i++) {
    const float cur = c.useAbsVal ? fabsf(x[i]) : x[i];
    if (cur > c.thresholdOnset && oo == 0) { nOnsets++; if (onsetPos == -1) onsetPos = i; oo = 1; }
    if (cur <= c.thresholdOffset && oo == 1) { nOffsets++; offsetPos = i; oo = 0; }
  }
  if (offsetPos == -1) offsetPos = N - 1;
  if (onsetPos == -1) onsetPos = 0;
  int n = 0;
  if (timeNorm == OSM_B200_TIMENORM_SEGMENT) {
    if (c.onsetPos) out[n++] = (float)onsetPos / (float)N;
    if (c.offsetPos) out[n++] = (float)offsetPos / (float)N;
  } else if (timeNorm == OSM_B200_TIMENORM_SECOND) {
    if (c.onsetPos) out[n++] = (float)onsetPos * period;
    if (c.offsetPos) out[n++] = (float)offsetPos * period;
  } else {
    if (c.onsetPos) out[n++] = (float)onsetPos;
    if (c.offsetPos) out[n++] = (float)offsetPos;
  }
  if (c.numOnsets) out[n++] = (float)nOnsets;
  if (c.numOffsets) out[n++] = (float)nOffsets;
  if (c.onsetRate) out[n++] = (float)nOnsets / ((float)N * period);
  return n;
}

// cFunctionalPeaks::process (functionalPeaks.cpp:96-213) with overlapFlag = 1 (its default: the two-sample history restarts with
// every contour).  dists: work space for the peak distances (at most N / 2 entries)
template <class Spec>
OSM_FS_HD int peaks(const Spec &c, const float *x, long N, float period, int timeNorm, int *dists, float *out)
{
  float mxv = x[0], mnv = x[0], mean = x[0];
  for (long i = 1; i < N; i++) { if (x[i] < mnv) mnv = x[i]; if (x[i] > mxv) mxv = x[i]; mean = mean + x[i]; }
  mean = mean / (float)N;
  const float range = mxv - mnv;
  float peakDist = 0.0f, peakMean = 0.0f, lastMin = 0.0f, lastMax = 0.0f;
  long nPeakDist = 0, nPeaks = 0, curmaxPos = 0, lastmaxPos = -1;
  int peakflag = 0;
  float lastlastVal = x[0], lastVal = N > 1 ? x[1] : 0.0f;
  for (long i = 2; i < N; i++) {
    if (lastlastVal < lastVal && lastVal > x[i]) {                                        // max
      if (!peakflag) lastMax = x[i];
      else if (x[i] > lastMax) { lastMax = x[i]; curmaxPos = i; }
      if ((double)(lastMax - lastMin) > 0.11 * (double)range) { peakflag = 1; curmaxPos = i; }
    } else if (lastlastVal > lastVal && lastVal < x[i]) lastMin = x[i];                   // min
    if (peakflag && ((double)x[i] < (double)lastMax - 0.09 * (double)range || i == N - 1)) {
      nPeaks++;
      peakMean = peakMean + lastMax;
      if (lastmaxPos >= 0) {
        const float dist = (float)(curmaxPos - lastmaxPos);
        peakDist = peakDist + dist;
        dists[nPeakDist++] = (int)dist;
      }
      lastmaxPos = curmaxPos;
      peakflag = 0;
    }
    lastlastVal = lastVal;
    lastVal = x[i];
  }
  float stddev = 0.0f;
  if (nPeakDist > 0) {
    peakDist = peakDist / (float)nPeakDist;
    for (long i = 0; i < nPeakDist; i++) stddev = stddev + ((float)dists[i] - peakDist) * ((float)dists[i] - peakDist);
    stddev = stddev / (float)nPeakDist;
    stddev = sqrtf(stddev);
  } else { peakDist = (float)(N + 1); stddev = 0.0f; }
  int n = 0;
  if (c.numPeaks) out[n++] = (float)nPeaks;
  if (timeNorm == OSM_B200_TIMENORM_SECOND) { peakDist = peakDist * period; stddev = stddev * period; }
  else if (timeNorm == OSM_B200_TIMENORM_SEGMENT) { peakDist = peakDist / (float)N; stddev = stddev / (float)N; }
  if (c.meanPeakDist) out[n++] = peakDist;
  peakMean = nPeaks > 0 ? peakMean / (float)nPeaks : 0.0f;
  if (c.peakMean) out[n++] = peakMean;
  if (c.peakMeanMeanDist) out[n++] = peakMean - mean;
  if (c.peakDistStddev) out[n++] = stddev;
  return n;
}

// cFunctionalCrossings::process (functionalCrossings.cpp:64-97): products in float, the mean-crossing terms in double
template <class Spec>
OSM_FS_HD int crossings(const Spec &c, const float *x, long N, float *out)
{
  double amean = 0.0;
  if (c.mcr || c.amean) {
    amean = (double)x[0];
    for (long i = 1; i < N; i++) amean += (double)x[i];
    amean /= (double)N;
  }
  long zcr = 0, mcr = 0;
  for (long i = 1; i < N - 1; i++) {
    const float a = x[i - 1], b = x[i], d = x[i + 1];
    if ((a * d <= 0.0f && b == 0.0f) || a * b < 0.0f) zcr++;
    if (c.mcr) {
      const double am = (double)a - amean, bm = (double)b - amean, dm = (double)d - amean;
      if ((am * dm <= 0.0 && bm == 0.0) || am * bm < 0.0) mcr++;
    }
  }
  int n = 0;
  if (c.zcr) out[n++] = (float)((double)zcr / (double)N);
  if (c.mcr) out[n++] = (float)((double)mcr / (double)N);
  if (c.amean) out[n++] = (float)amean;
  return n;
}

// cFunctionalSamples::process (functionalSamples.cpp:99-116): the contour's value at relative positions
template <class Spec>
OSM_FS_HD int samples(const Spec &c, const float *x, long N, float *out)
{
  const float Nind = (float)N;
  for (int k = 0; k < c.n_samplepos; k++) out[k] = x[(int)(((double)Nind - 1.0) * c.samplepos[k])];
  return c.n_samplepos;
}

// cFunctionalDCT::process (functionalDCT.cpp:85-135), coefficient `coeff` (absolute index): the table entry is
// (float)cos(pi * i / N * ((float)m + 0.5)) in double, the sum runs in float over m = 0 .. N-1, times (float)sqrt(2 / N)
OSM_FS_HD float dct_coeff(const float *x, long N, int coeff)
{
  float acc = 0.0f;
  const double w = M_PI * (double)coeff / (double)N;
  for (long m = 0; m < N; m++) {
    const float ct = (float)cos(w * ((double)(float)m + 0.5));
    const float pr = x[m] * ct;
    acc = acc + pr;
  }
  acc = acc * (float)sqrt(2.0 / (double)N);
  return isfinite(acc) ? acc : 0.0f;
}

}  // namespace fseq
}  // namespace osm
