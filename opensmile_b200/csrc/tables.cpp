// tables.cpp -- host-side construction of the constant tables the fused kernel consumes.
// These are the "finalise time" computations of the reference's components
// (cWindower::precomputeWinFunc, cMelspec::computeFilters, cMfcc::initTables); they run once
// per plan on the CPU exactly as the reference runs them once per component instance, with
// the same float/double casting order so that the tables are bit-identical.
// Citations are relative to /root/reference/src.
#include <cmath>
#include <cstring>

#include "plan.hpp"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace osm {

// cWindower::precomputeWinFunc (dspcore/windower.cpp:159-217) over the double tables of
// smileutil/smileUtil.c:1218-1349; the per-sample use casts to float (windower.cpp:226).
void build_window(int winFunc, int N, double sigma, double gain, std::vector<float> &out)
{
  std::vector<double> w(N);
  const double NN = (double)N;
  for (int n = 0; n < N; n++) {
    const double i = (double)n;
    switch (winFunc) {
      case OSM_B200_WIN_HANNING:   // smileUtil.c:1277-1288
        w[n] = 0.5 * (1.0 - cos((2.0 * M_PI * i) / (NN - 1.0)));
        break;
      case OSM_B200_WIN_HAMMING:   // smileUtil.c:1291-1303
        w[n] = 0.54 - 0.46 * cos((2.0 * M_PI * i) / (NN - 1.0));
        break;
      case OSM_B200_WIN_GAUSS: {   // smileUtil.c:1334-1349
        double s = sigma;
        if (s <= 0.0) s = 0.01;
        if (s > 0.5) s = 0.5;
        const double tmp = (i - (NN - 1.0) / 2.0) / (s * (NN - 1.0) / 2.0);
        w[n] = exp(-0.5 * (tmp * tmp));
        break;
      }
      case OSM_B200_WIN_SINE:      // smileUtil.c:1306-1317
        w[n] = sin((1.0 * M_PI * i) / (NN - 1.0));
        break;
      case OSM_B200_WIN_TRIANGLE:  // smileUtil.c:1232-1246
        w[n] = (n < N / 2) ? 2.0 * (double)(n + 1) / (double)N : 2.0 * (double)(N - n) / (double)N;
        break;
      case OSM_B200_WIN_BARTLETT:  // smileUtil.c:1261-1274
        w[n] = (n < N / 2) ? 2.0 * (double)n / (double)(N - 1)
                           : 2.0 * (double)(N - 1 - n) / (double)(N - 1);
        break;
      default:                     // rectangle, smileUtil.c:1218-1228
        w[n] = 1.0;
        break;
    }
  }
  if (gain != 1.0)
    for (int n = 0; n < N; n++) w[n] *= gain;  // windower.cpp:192-196
  out.resize(N);
  for (int n = 0; n < N; n++) out[n] = (float)w[n];
}

// smileDsp_specScaleTransfFwd / Inv for SPECTSCALE_MEL (smileutil/smileUtil.c:1139-1142,1197)
static double mel_fwd(double x) { return x > 0.0 ? 1127.0 * log(1.0 + x / 700.0) : 0.0; }
static double mel_inv(double x) { return 700.0 * (exp(x / 1127.0) - 1.0); }

// cMelspec::computeFilters, standard triangular bank (lldcore/melspec.cpp:184-240,391-447),
// specScale = mel (forced when htkcompatible, melspec.cpp:127-131).
void build_mel(const osm_b200_melspec &cfg, int blocksize, double frameSizeSec, MelBank &mb)
{
  const int nBands = cfg.nBands;
  mb.nBands = nBands;
  mb.nBins = blocksize;
  mb.coef.assign(blocksize, 0.f);
  mb.chanMap.assign(blocksize, -3);
  mb.bandHz.assign(nBands, 0.0);
  std::vector<float> cfs(nBands + 2);

  const float N = (float)((blocksize - 1) * 2);               // :217
  const float F0 = (float)(1.0 / frameSizeSec);               // :220
  const float Fs = (float)(N / frameSizeSec);                 // :221
  const float M = (float)nBands;
  float lofreq = (float)cfg.lofreq, hifreq = (float)cfg.hifreq;  // FLOAT_DMEM members, melspec.hpp:48
  if ((lofreq < 0.0) || (lofreq > Fs / 2.0) || (lofreq > hifreq)) lofreq = 0.0;             // :224
  if ((hifreq < lofreq) || (hifreq > Fs / 2.0) || (hifreq <= 0.0)) hifreq = Fs / (float)2.0; // :226
  const float LoF = (float)mel_fwd(lofreq);                   // :228
  const float HiF = (float)mel_fwd(hifreq);                   // :230
  long nLoF = (long)round((double)(lofreq / F0));             // FtoN, melspec.hpp:107-110
  long nHiF = (long)round((double)(hifreq / F0));
  if (nLoF > blocksize) nLoF = blocksize;
  if (nHiF > blocksize) nHiF = blocksize;
  if (nLoF < 0) nLoF = 0;
  if (nHiF < 0) nHiF = 0;
  mb.nLo = (int)nLoF;
  mb.nHi = (int)nHiF;

  const float mBandw = (HiF - LoF) / (M + (float)1.0);        // :394
  for (int m = 0; m <= nBands + 1; m++) cfs[m] = LoF + (float)m * mBandw;   // :395-397
  for (int m = 1; m <= nBands; m++) mb.bandHz[m - 1] = mel_inv(cfs[m]);     // :408-411

  // channel map :427-438 ; NtoFmel(n,F0) = (float)fwd((float)n * F0), melspec.hpp:119-122
  int m = 0;
  for (int n = 0; n < blocksize; n++) {
    if ((n <= nLoF) || (n >= nHiF)) {
      mb.chanMap[n] = -3;
    } else {
      while (cfs[m] < (float)mel_fwd(((float)n) * F0)) {
        if (m > nBands) break;
        m++;
      }
      mb.chanMap[n] = m - 2;
    }
  }
  // rising-slope weights :441-447
  m = 0;
  for (long n = nLoF; n < nHiF; n++) {
    const float nM = (float)mel_fwd(((float)n) * F0);
    while ((nM > cfs[m + 1]) && (m <= nBands)) m++;
    mb.coef[n] = (cfs[m + 1] - nM) / (cfs[m + 1] - cfs[m]);
  }

  // kernel view: the visited bins nLo..nHi-1 with chanMap > -2 form nBands+1 contiguous
  // runs ("ranges"), range r holding the bins with chanMap == r-1.  The per-frame loop
  // (melspec.cpp:543-553) then is, for a bin of range r:  band[r-1] += a ; band[r] += p - a.
  mb.rangeBegin.assign(nBands + 2, 0);
  {
    int n = (int)nLoF;
    while (n < nHiF && mb.chanMap[n] <= -2) n++;  // skipped bins (chanMap -3) at the low edge
    for (int r = 0; r <= nBands; r++) {
      mb.rangeBegin[r] = n;
      while (n < nHiF && mb.chanMap[n] == r - 1) n++;
    }
    mb.rangeBegin[nBands + 1] = n;
    // anything left (chanMap >= nBands or -3 inside) contributes nothing in the reference
    // either only if chanMap <= -2; a chanMap >= nBands cannot occur (m-2 <= nBands-1).
  }
  mb.usePower = cfg.usePower != 0;
  if (cfg.htkcompatible)                                       // :559-569
    mb.outScale = cfg.usePower ? (float)(32767.0 * 32767.0) : (float)32767.0;
  else
    mb.outScale = 1.f;
}

// cMfcc::initTables (lldcore/mfcc.cpp:136-170) + the output-order permutation and lifter
// product of processVector (:251-272).
void build_mfcc(const osm_b200_mfcc &cfg, int nBands, MfccOp &op)
{
  const int first = cfg.firstMfcc, last = cfg.lastMfcc;
  const int nM = last - first + 1;
  op.first = first; op.last = last; op.nMfcc = nM;
  float melfloor = (float)cfg.melfloor;                        // :71
  if (cfg.htkcompatible) melfloor = 1.0f;                      // :88-91
  op.melfloor = melfloor;
  op.logMelfloor = std::log(melfloor);                         // float overload, :240
  op.doLog = cfg.doLog != 0;
  const float cepLifter = (float)cfg.cepLifter;                // :75

  std::vector<float> cost((size_t)nBands * nM), sint(nM);
  const double fnM = (double)nBands;
  for (int i = first; i <= last; i++) {                        // :146-152
    const double fi = (double)i;
    for (int m = 0; m < nBands; m++)
      cost[m + (i - first) * nBands] = (float)cos((double)M_PI * (fi / fnM) * ((double)m + 0.5));
  }
  for (int i = first; i <= last; i++) {                        // :158-166
    if (cepLifter > 0.0)
      sint[i - first] = ((float)1.0 + cepLifter / (float)2.0 * std::sin((float)M_PI * ((float)i) / cepLifter));
    else
      sint[i - first] = 1.0f;
  }
  const float factor = (float)sqrt((double)2.0 / (double)nBands);  // :251
  op.cosT.assign((size_t)nBands * nM, 0.f);
  op.liftFactor.assign(nM, 0.f);
  for (int i = first; i <= last; i++) {                        // :252-258 output slot -> table row
    const int slot = i - first;
    int i0 = slot;
    if (cfg.htkcompatible && first == 0) i0 = (i == last) ? 0 : i0 + 1;
    memcpy(&op.cosT[(size_t)slot * nBands], &cost[(size_t)i0 * nBands], sizeof(float) * nBands);
    op.liftFactor[slot] = sint[i0] * factor;                   // :272
  }
}

// smileDsp_equalLoudnessWeight(_htk), smileutil/smileUtil.c:1041-1061
static double eql_htk(double f)
{
  const double f2 = f * f, fs = f2 / (f2 + 1.6e5);
  return fs * fs * ((f2 + 1.44e6) / (f2 + 9.61e6));
}
static double eql_hermansky(double f)
{
  const double w = 2.0 * M_PI * f, w2 = w * w, c = w2 + 6300000.0;
  if (c > 0.0) return (1e32 * ((w2 + 56.8e6) * w2 * w2) / (c * c * (w2 + 0.38e9) * (w2 * w2 * w2 * w + 1.7e31)));
  return 0.0;
}

// cPlp::myFetchConfig (lldcore/plp.cpp:88-171) + cPlp::initTables (:276-341).
// RASTA / newRASTA (a recurrence over frames) is not fused yet.
bool build_plp(const osm_b200_plp &cfg, const MelBank &mb, PlpOp &op, std::string &err)
{
  if (cfg.RASTA || cfg.newRASTA) { err = "cPlp: RASTA / newRASTA are not supported yet"; return false; }
  int lpOrder = cfg.lpOrder;
  bool doLP = cfg.doLP != 0, doLpToCeps = cfg.doLpToCeps != 0, doIDFT = cfg.doIDFT != 0;
  if (lpOrder <= 0) { lpOrder = 0; doLP = false; doLpToCeps = false; }                 // :103-106
  int nCeps = cfg.nCeps, firstCC = cfg.firstCC, lastCC = cfg.lastCC;
  if (firstCC > lpOrder) { firstCC = lpOrder; nCeps = 1; lastCC = lpOrder; }           // :111-112
  else if (firstCC < 0) firstCC = 0;
  if (nCeps < 0) nCeps = lpOrder - firstCC + 1;                                         // :114-116
  if (lastCC < 0) lastCC = firstCC + nCeps - 1;                                         // :118-121
  else if (lastCC >= firstCC) nCeps = lastCC - firstCC + 1;
  if (lastCC > lpOrder) { lastCC = lpOrder; nCeps = lastCC - firstCC + 1; }             // :122-126
  if (nCeps == 0) doLpToCeps = false;                                                   // :132
  if (doLpToCeps) doLP = true;                                                          // :134-136
  if (doLP) doIDFT = true;                                                              // :137-139
  if (lpOrder > 8) { err = "cPlp: lpOrder > 8 is not supported"; return false; }
  op.lpOrder = lpOrder; op.nCeps = nCeps; op.firstCC = firstCC; op.lastCC = lastCC;
  op.doLP = doLP; op.doLpToCeps = doLpToCeps; op.doIDFT = doIDFT;
  float compression = (float)cfg.compression;                                           // :142-143
  if (compression < 0.0) compression = 0.0;
  op.compression = compression;
  float cepLifter = (float)(int)cfg.cepLifter;                                          // :146 getInt
  if (cepLifter < 0) cepLifter = 0;
  op.melfloor = (float)cfg.melfloor;
  op.doLog = cfg.doLog != 0; op.doAud = cfg.doAud != 0; op.doInvLog = cfg.doInvLog != 0;
  op.htk = cfg.htkcompatible != 0;
  if (op.htk) { op.melfloor = 1.0f; op.doAud = true; op.doLog = false; op.doInvLog = false; }   // :152-163
  op.logMelfloor = std::log(op.melfloor);
  const int nBands = mb.nBands;
  op.nFreq = nBands + 2;                                                                // :288
  op.nAuto = lpOrder + 1;
  op.cosT.assign((size_t)op.nAuto * op.nFreq, 0.f);
  {
    const float a = (float)M_PI / (float)(op.nFreq - 1);                                // :298
    for (int i = 0; i < op.nAuto; i++) {
      const int ib = i * op.nFreq;
      int m;
      op.cosT[ib] = 1.0;
      for (m = 1; m < (op.nFreq - 1); m++) op.cosT[m + ib] = (float)(2.0 * cos(a * (double)i * (double)m));
      op.cosT[m + ib] = (float)(cos(a * (double)i * (double)m));
    }
  }
  op.lifter = cepLifter > 0.0;
  std::vector<float> sint(nCeps > 0 ? nCeps : 1, 1.f);
  for (int i = firstCC; i <= lastCC; i++) {                                             // :320-327
    if (cepLifter > 0.0)
      sint[i - firstCC] = ((float)1.0 + cepLifter / (float)2.0 * std::sin((float)M_PI * ((float)(i)) / cepLifter));
    else
      sint[i - firstCC] = 1.0;
  }
  op.lift.assign(nCeps > 0 ? nCeps : 1, 1.f);
  for (int i = firstCC; i <= lastCC; i++) {                                             // :560-573 output slot -> table index
    const int i0 = i - firstCC;
    int i1 = i0;
    if (op.htk && firstCC == 0) i1 = (i == lastCC) ? 0 : i1 + 1;
    op.lift[i0] = sint[i1];
  }
  op.eql.assign(nBands, 1.f);
  for (int i = 0; i < nBands; i++) {                                                    // :345-357 (band centres = melspec field info)
    op.eql[i] = op.htk ? (float)eql_htk(mb.bandHz[i]) : (float)eql_hermansky(mb.bandHz[i]);
    if (op.doLog) op.eql[i] = std::log(op.eql[i]);
  }
  op.nOut = doLpToCeps ? nCeps : (doLP ? lpOrder : (doIDFT ? op.nAuto : nBands));        // :232-267
  return true;
}

}  // namespace osm
