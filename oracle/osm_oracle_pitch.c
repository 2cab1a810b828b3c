/*
 * osm_oracle_pitch.c -- CPU restatement of cSpecScale -> cPitchShs -> cPitchSmootherViterbi ->
 * cValbasedSelector -> cPitchJitter.  TEST INFRASTRUCTURE ONLY, see
 * osm_oracle_pitch.h.  Citations relative to /root/reference/src.  Arithmetic types follow the
 * reference (FLOAT_DMEM = float; double where the reference computes in double).
 */
#include "osm_oracle_pitch.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* smileutil/smileUtil.c:720-723 */
static double log2_ref(double x) { return log(x) / log(2.0); }

/* smileutil/smileUtil.c:1009-1034: vertex of the parabola through three points */
static double quad3(double x1, double y1, double x2, double y2, double x3, double y3, double *y, double *a_out)
{
  double den = x1 * x1 * x2 + x2 * x2 * x3 + x3 * x3 * x1 - x3 * x3 * x2 - x2 * x2 * x1 - x1 * x1 * x3;
  if (den != 0.0) {
    double a = (y1 * x2 + y2 * x3 + y3 * x1 - y3 * x2 - y2 * x1 - y1 * x3) / den;
    double b = (x1 * x1 * y2 + x2 * x2 * y3 + x3 * x3 * y1 - x3 * x3 * y2 - x2 * x2 * y1 - x1 * x1 * y3) / den;
    double c = (x1 * x1 * x2 * y3 + x2 * x2 * x3 * y1 + x3 * x3 * x1 * y2 - x3 * x3 * x2 * y1 - x2 * x2 * x1 * y3 - x1 * x1 * x3 * y2) / den;
    if (a != 0.0) {
      if (a_out) *a_out = a;
      double x = -b / (2.0 * a);
      if (y) *y = c - a * x * x;
      return x;
    }
  }
  if (a_out) *a_out = 0.0;
  if (y1 > y2 && y1 > y3) { if (y) *y = y1; return x1; }
  else if (y2 > y1 && y2 > y3) { if (y) *y = y2; return x2; }
  else if (y3 > y1 && y3 > y2) { if (y) *y = y3; return x3; }
  if (y) *y = y1;
  return x1;
}

/* ------------------------------------------------------------------ cSpecScale */

typedef struct {
  long nMag, nPts;
  double *f_t;                 /* octave position of every source bin (specScale.cpp:262-268) */
  double *sigma, *diff1, *diff2;   /* spline cache (smileUtilSpline.c:124-140) */
  long *k; double *coef;       /* interpolation cache (:301-352) */
  double *audw;                /* auditory weighting (specScale.cpp:289-297) */
  double *y, *y2, *u;
  float nOctaves, nPointsPerOctave, fmin_t, fmax_t, minF;   /* level meta data as floats (:299-311) */
  int smooth, enhance;
} scale_ctx;

static int scale_init(scale_ctx *s, const osm_or_specscale_cfg *sc, long nMag, double fsSecLevel)
{
  memset(s, 0, sizeof *s);
  s->nMag = nMag;
  s->nPts = sc->nPointsTarget > 0 ? sc->nPointsTarget : nMag;          /* specScale.cpp:205-207 */
  s->smooth = sc->specSmooth; s->enhance = sc->specEnhance;
  double fsSec = (double)(float)fsSecLevel;                             /* :184-187 */
  double deltaF = 1.0 / fsSec;                                          /* :204 */
  double minF = sc->minF < 1.0 ? 1.0 : sc->minF, maxF = sc->maxF;       /* :160-164 */
  double samplF = deltaF * (double)(nMag - 1);                          /* :244-247 */
  if (maxF <= minF || maxF > samplF) maxF = samplF;
  double fmin_t = log(minF) / log(2.0), fmax_t = log(maxF) / log(2.0);  /* :251-252, smileUtil.c:1100-1102 */
  double deltaF_t = (fmax_t - fmin_t) / (double)(s->nPts - 1);          /* :255 */
  s->f_t = (double *)malloc(sizeof(double) * nMag);
  for (long i = 1; i < nMag; i++) s->f_t[i] = log((double)i * deltaF) / log(2.0);   /* :260-263 */
  s->f_t[0] = 2.0 * s->f_t[1] - s->f_t[2];                              /* :264 */
  s->sigma = (double *)calloc(nMag, sizeof(double));
  s->diff1 = (double *)calloc(nMag, sizeof(double));
  s->diff2 = (double *)calloc(nMag, sizeof(double));
  const double *x = s->f_t;
  for (long i = 1; i < nMag - 1; i++) {                                 /* smileUtilSpline.c:130-134 */
    s->sigma[i] = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    s->diff1[i] = (x[i + 1] - x[i]) * (x[i + 1] - x[i - 1]);
    s->diff2[i] = (x[i] - x[i - 1]) * (x[i + 1] - x[i - 1]);
  }
  s->k = (long *)malloc(sizeof(long) * s->nPts);
  s->coef = (double *)malloc(sizeof(double) * 3 * s->nPts);
  long kupper = 1;
  for (long i = 0; i < s->nPts; i++) {                                  /* smileUtilSpline.c:301-352 */
    double xi = fmin_t + (double)i * deltaF_t;                          /* specScale.cpp:276-278 */
    if (i == 0 && xi < x[0]) return 0;
    while (kupper < nMag && x[kupper] < xi) kupper++;
    if (kupper == nMag) return 0;                                       /* "x out of range": output invalid */
    long klower = kupper - 1;
    s->k[i] = klower;
    double range = x[kupper] - x[klower];
    if (range == 0.0) return 0;
    double a = (x[kupper] - xi) / range, b = 1.0 - a, range2 = range * range / 6.0;
    s->coef[i * 3] = a;
    s->coef[i * 3 + 1] = (a * a * a - a) * range2;
    s->coef[i * 3 + 2] = (b * b * b - b) * range2;
  }
  double nOct = log(maxF / minF) / log(2.0);                            /* :286-287 */
  double ppo = (double)s->nPts / nOct;
  if (sc->auditoryWeighting) {
    double atan_s = ppo * log2_ref(65.0 / 50.0) - 1.0;                  /* :290-296 */
    s->audw = (double *)malloc(sizeof(double) * s->nPts);
    for (long i = 0; i < s->nPts; i++) s->audw[i] = 0.5 + atan(3.0 * ((double)(i + 1) - atan_s) / ppo) / M_PI;
  }
  s->minF = (float)minF; s->nOctaves = (float)nOct; s->nPointsPerOctave = (float)ppo;
  s->fmin_t = (float)fmin_t; s->fmax_t = (float)fmax_t;
  s->y = (double *)malloc(sizeof(double) * nMag);
  s->y2 = (double *)malloc(sizeof(double) * nMag);
  s->u = (double *)calloc(nMag, sizeof(double));
  return 1;
}

static void scale_free(scale_ctx *s)
{
  free(s->f_t); free(s->sigma); free(s->diff1); free(s->diff2); free(s->k); free(s->coef);
  free(s->audw); free(s->y); free(s->y2); free(s->u);
}

/* smileutil/smileUtil.c:1965-2003: zero everything that is further than 2 bins from a local maximum */
static void spec_enhance(double *a, long n)
{
  if (n < 2) return;
  long *posmax = (long *)calloc((size_t)((n + 1) / 2 + 1), sizeof(long));
  long nmax = 0;
  if (a[0] > a[1]) posmax[nmax++] = 0;
  for (long i = 1; i < n - 1; i++) if (a[i] > a[i - 1] && a[i] >= a[i + 1]) posmax[nmax++] = i;
  if (a[n - 1] > a[n - 2]) posmax[nmax++] = n - 1;
  if (nmax == 1) {                       /* reads posmax[1] (== 0, calloc) like the reference */
    for (long j = 0; j <= posmax[1] - 3; j++) a[j] = 0;
    for (long j = posmax[1] + 3; j < n; j++) a[j] = 0;
  } else {
    for (long i = 1; i < nmax; i++)
      for (long j = posmax[i - 1] + 3; j <= posmax[i] - 3; j++) a[j] = 0;
  }
  free(posmax);
}

/* smileutil/smileUtil.c:2006-2016 */
static void spec_smooth(double *a, long n)
{
  double aim1 = 0.0;
  for (long i = 0; i < n - 1; i++) {
    double ai = a[i];
    a[i] = (aim1 + 2.0 * ai + a[i + 1]) / 4.0;
    aim1 = ai;
  }
}

/* dsp/specScale.cpp:318-371 */
static void scale_frame(scale_ctx *s, const float *mag, float *dst)
{
  long N = s->nMag;
  double *y = s->y, *y2 = s->y2, *u = s->u;
  for (long i = 0; i < N; i++) y[i] = (double)mag[i];
  if (s->enhance) spec_enhance(y, N);
  if (s->smooth) spec_smooth(y, N);
  /* natural cubic spline, cached abscissa terms (smileUtilSpline.c:142-190) */
  u[0] = 0.0; y2[0] = 0.0;
  for (long i = 1; i < N - 1; i++) {
    double sg = s->sigma[i];
    double p = 1.0 / (sg * y2[i - 1] + 2.0);
    y2[i] = (sg - 1.0) * p;
    double ut = (y[i + 1] - y[i]) / s->diff1[i] - (y[i] - y[i - 1]) / s->diff2[i];
    u[i] = p * (6.0 * ut - sg * u[i - 1]);
  }
  y2[N - 1] = (0.0 - 0.0 * u[N - 2]) / (0.0 * y2[N - 2] + 1.0);
  for (long j = N - 2; j >= 0; j--) y2[j] = y2[j] * y2[j + 1] + u[j];
  for (long i = 0; i < s->nPts; i++) {                                  /* smileUtilSpline.c:355-368 */
    double a = s->coef[i * 3], b = 1.0 - a, c = s->coef[i * 3 + 1], d = s->coef[i * 3 + 2];
    long k = s->k[i];
    dst[i] = (float)(a * y[k] + b * y[k + 1] + c * y2[k] + d * y2[k + 1]);
  }
  if (s->audw) {                                                        /* specScale.cpp:360-368 */
    for (long i = 0; i < s->nPts; i++) {
      if (dst[i] > 0.0f) dst[i] = (float)((double)dst[i] * s->audw[i]);
      else dst[i] = 0.0f;
    }
  }
}

/* ------------------------------------------------------------------ cPitchShs / cPitchBase */

int osm_or_pitchshs_num_out(const osm_or_pitchshs_cfg *ps)
{
  int nc = ps->nCandidates < 1 ? 1 : (ps->nCandidates > 20 ? 20 : ps->nCandidates);   /* pitchBase.cpp:88-90 */
  return 1 + nc + (ps->voicing ? nc : 0) + (ps->scores ? nc : 0) + (ps->F0C1 ? 1 : 0) + (ps->voicingC1 ? 1 : 0) +
         (ps->F0raw ? 1 : 0) + (ps->voicingClip ? 1 : 0);
}

typedef struct {
  const osm_or_pitchshs_cfg *ps;
  int nc;
  long N;
  float Fmint, Fstept, nOctaves, nPointsPerOctave, compression, voicingCutoff;
  double base, maxPitch, minPitch;
  float *SS, *in;
} shs_ctx;

/* lld/pitchShs.cpp:220-358 */
static int shs_detect(shs_ctx *c, float *inData, float *f0cand, float *candVoice, float *candScore)
{
  const osm_or_pitchshs_cfg *ps = c->ps;
  long N = c->N, i, j;
  int nCand = 0, nC = c->nc;
  float *SS = c->SS;
  if (c->nOctaves == 0.0f) return -1;
  if (ps->lfCut > 0.0) {                                                /* :230-236 */
    int bin = (int)((ceil(log(ps->lfCut) / log(c->base)) - c->Fmint) / c->Fstept);
    for (i = 0; i <= bin; i++) inData[i] = 0.0f;
  }
  for (j = 0; j < N; j++) SS[j] = inData[j];
  float scale = c->compression;
  for (i = 2; i < ps->nHarmonics + 1; i++) {                            /* :248-254 */
    long shift = (long)floor((double)c->nPointsPerOctave * log2_ref((double)i));
    for (j = shift; j < N; j++) SS[j - shift] += inData[j] * scale;
    scale *= c->compression;
  }
  for (j = 0; j < N; j++) {
    SS[j] /= (float)ps->nHarmonics;
    if (SS[j] < 0) SS[j] = 0.0f;
  }
  candScore[0] = 0.0f;
  double ssMean = (double)SS[0];
  for (i = 1; i < N - 1; i++) {                                         /* :271-320 */
    if (ps->greedyPeakAlgo) {
      if (SS[i - 1] < SS[i] && SS[i] > SS[i + 1]) {
        for (j = 0; j < nC; j++) {
          if (candScore[j] == 0.0f || candScore[j] < SS[i]) {
            for (long jj = nC - 1; jj > j; jj--) { candScore[jj] = candScore[jj - 1]; f0cand[jj] = f0cand[jj - 1]; }
            f0cand[j] = (float)i;
            candScore[j] = SS[i];
            if (nCand < nC) nCand++;
            break;
          }
        }
      }
    } else {
      if (SS[i - 1] < SS[i] && SS[i] > SS[i + 1] && (SS[i] > candScore[0] || candScore[0] == 0.0f)) {
        for (j = nC - 1; j > 0; j--) { candScore[j] = candScore[j - 1]; f0cand[j] = f0cand[j - 1]; }
        f0cand[0] = (float)i;
        candScore[0] = SS[i];
        if (nCand < nC) nCand++;
      }
    }
    ssMean += (double)SS[i];
  }
  ssMean = (ssMean + (double)SS[i]) / (double)N;
  for (i = 0; i < nCand; i++) {                                         /* :323-343 */
    long jx = (long)f0cand[i];
    float f1 = f0cand[i] * c->Fstept + c->Fmint;
    float f2 = (f0cand[i] + (float)1.0) * c->Fstept + c->Fmint;
    float f0 = (f0cand[i] - (float)1.0) * c->Fstept + c->Fmint;
    double sc = 0;
    double fx = quad3((double)f0, (double)SS[jx - 1], (double)f1, (double)SS[jx], (double)f2, (double)SS[jx + 1], &sc, NULL);
    f0cand[i] = (float)exp(fx * log(c->base));
    candScore[i] = (float)sc;
    if (sc > 0.0 && sc > ssMean) candVoice[i] = (float)(1.0 - ssMean / sc);
    else candVoice[i] = 0.0f;
  }
  if (ps->octaveCorrection) {                                           /* :346-358 */
    for (i = 1; i < nCand; i++) {
      if (f0cand[i] < f0cand[0] && f0cand[i] > 0 &&
          (candVoice[i] > c->voicingCutoff || candVoice[i] >= 0.9 * c->voicingCutoff) &&
          candScore[i] > ((1.0 / (float)(ps->nHarmonics - 1) * c->compression)) * candScore[0]) {
        float t;
        t = f0cand[0]; f0cand[0] = f0cand[i]; f0cand[i] = t;
        t = candVoice[0]; candVoice[0] = candVoice[i]; candVoice[i] = t;
        t = candScore[0]; candScore[0] = candScore[i]; candScore[i] = t;
      }
    }
  }
  return nCand;
}

/* lldcore/pitchBase.cpp:173-300 */
static void shs_frame(shs_ctx *c, const float *hps, float *dst)
{
  const osm_or_pitchshs_cfg *ps = c->ps;
  int nC = c->nc;
  float f0cand[20], candVoice[20], candScore[20];
  long i, j;
  for (i = 0; i < nC; i++) { f0cand[i] = 0.0f; candVoice[i] = 0.0f; candScore[i] = 0.0f; }
  memcpy(c->in, hps, sizeof(float) * c->N);
  int nCand = shs_detect(c, c->in, f0cand, candVoice, candScore);
  if (nCand > 0) {                                                      /* :196-213 */
    for (i = 0; i < nC && nCand > 0; i++) {
      if ((double)f0cand[i] > c->maxPitch || (double)f0cand[i] < c->minPitch) {
        float origF = f0cand[i];
        for (j = i + 1; j < nC; j++) { f0cand[j - 1] = f0cand[j]; candVoice[j - 1] = candVoice[j]; candScore[j - 1] = candScore[j]; }
        f0cand[j - 1] = 0; candVoice[j - 1] = 0; candScore[j - 1] = 0;
        if (origF > 0.0f) { nCand--; i--; }
      }
    }
  }
  int n = 0;
  if (nCand < 0) { int K = osm_or_pitchshs_num_out(ps); for (i = 0; i < K; i++) dst[i] = 0.0f; return; }
  dst[n++] = (float)nCand;
  long maxI = 0;
  if (!ps->octaveCorrection) {                                          /* :224-230 */
    float mx = candScore[0];
    for (i = 1; i < nC; i++) if (candScore[i] > mx) { mx = candScore[i]; maxI = i; }
  }
  if (maxI > 0) {
    float t;
    t = f0cand[0]; f0cand[0] = f0cand[maxI]; f0cand[maxI] = t;
    t = candVoice[0]; candVoice[0] = candVoice[maxI]; candVoice[maxI] = t;
    t = candScore[0]; candScore[0] = candScore[maxI]; candScore[maxI] = t;
  }
  for (i = 0; i < nC; i++) dst[n++] = f0cand[i];
  if (ps->voicing) for (i = 0; i < nC; i++) dst[n++] = candVoice[i];
  if (ps->scores) for (i = 0; i < nC; i++) dst[n++] = candScore[i];
  if (ps->F0C1) dst[n++] = f0cand[0];                                   /* :263-290 */
  if (ps->voicingC1) dst[n++] = candVoice[0];
  if (ps->F0raw) dst[n++] = candVoice[0] <= c->voicingCutoff ? 0.0f : f0cand[0];
  if (ps->voicingClip) dst[n++] = candVoice[0] <= c->voicingCutoff ? 0.0f : candVoice[0];
}

long osm_or_pitch_shs(const osm_or_frontend *fe, const osm_or_specscale_cfg *sc, const osm_or_pitchshs_cfg *ps,
                      const int16_t *pcm, long L, int n_chan, float *out_shs, float *tap_hps)
{
  long N = osm_or_frame_size_samples(fe), H = osm_or_frame_step_samples(fe);
  long nfft = osm_or_fft_size(N), nb = nfft / 2 + 1;
  long T = osm_or_num_frames(L, N, H);
  if (T <= 0) return 0;
  scale_ctx s;
  if (!scale_init(&s, sc, nb, osm_or_fft_frame_size_sec(fe))) { scale_free(&s); return -1; }
  shs_ctx c; memset(&c, 0, sizeof c);
  c.ps = ps;
  c.nc = ps->nCandidates < 1 ? 1 : (ps->nCandidates > 20 ? 20 : ps->nCandidates);
  c.N = s.nPts;
  c.nOctaves = s.nOctaves; c.nPointsPerOctave = s.nPointsPerOctave;     /* pitchShs.cpp:168-176 */
  c.base = exp(log((double)s.minF) / (double)s.fmin_t);                 /* :184-191 */
  if (fabs(c.base - 2.0) < 0.00001) c.base = 2.0;
  c.Fmint = s.fmin_t;
  c.Fstept = (s.fmax_t - s.fmin_t) / (float)(c.N - 1);                  /* :193-194 */
  c.compression = (float)ps->compressionFactor;
  c.voicingCutoff = (float)ps->voicingCutoff;
  c.maxPitch = ps->maxPitch < 0.0 ? 0.0 : ps->maxPitch;                 /* pitchBase.cpp:80-86 */
  c.minPitch = ps->minPitch < 0.0 ? 0.0 : ps->minPitch;
  if (c.minPitch > c.maxPitch) c.minPitch = c.maxPitch;
  c.SS = (float *)malloc(sizeof(float) * c.N);
  c.in = (float *)malloc(sizeof(float) * c.N);
  int K = osm_or_pitchshs_num_out(ps);
  float *x = (float *)malloc(sizeof(float) * L);
  osm_or_pcm16_to_float(pcm, L, n_chan, x);
  double *win = (double *)malloc(sizeof(double) * N);
  osm_or_window_table(fe->win_func, N, fe->win_sigma, fe->win_gain, win);
  float *mag = (float *)malloc(sizeof(float) * nb);
  float *hps = (float *)malloc(sizeof(float) * s.nPts);
  for (long t = 0; t < T; t++) {
    osm_or_frame_to_mag(fe, x + t * H, N, nfft, win, NULL, mag);
    scale_frame(&s, mag, hps);
    if (tap_hps) memcpy(tap_hps + t * s.nPts, hps, sizeof(float) * s.nPts);
    shs_frame(&c, hps, out_shs + t * K);
  }
  free(x); free(win); free(mag); free(hps); free(c.SS); free(c.in);
  scale_free(&s);
  return T;
}

/* ------------------------------------------------------------------ cPitchSmootherViterbi */

int osm_or_viterbi_num_out(const osm_or_viterbi_cfg *vc)
{
  return (vc->F0final ? 1 : 0) + (vc->F0finalLog ? 1 : 0) + (vc->F0finalEnv ? 1 : 0) + (vc->F0finalEnvLog ? 1 : 0) +
         (vc->voicingFinalClipped ? 1 : 0) + (vc->voicingFinalUnclipped ? 1 : 0);
}

typedef struct {
  int nStates; long buflen; int frameSize;
  long wrIdx, rdIdx, pathIdx, convIdx;
  int pathBuf;
  float *buf, *prev;
  int *paths[2], *bestPath;
  double *pathCosts, *pathCostsNew, *pathCostsTemp;
  float voiceThresh;
  double wLocal, wTvv, wTvvd, wTvuv, wTuu, wThr, wRange, lastChange;
} vit;

/* include/lld/pitchSmootherViterbi.hpp:167-197 */
static double vit_fweight(float f)
{
  if (f > 0.0 && f < 100.0) return -(1.0 / 100.0) * f + 1.0;
  else if (f >= 100.0 && f < 350.0) return 0.0;
  else if (f >= 350.0 && f < 600.0) return ((f - 350.0) / 250.0);
  else if (f >= 600.0) return 1.2;
  else if (f <= 0) return 2.0;
  return 0.0;
}

/* :202-221 */
static double vit_local(vit *v, int i, const float *frame)
{
  double pv = (double)frame[i * 2 + 1];
  double thr = 0.0;
  if (pv < 0.01) pv = 0.01;
  if (pv > 1.00) pv = 1.00;
  if (pv < v->voiceThresh) thr = v->wThr;
  if (i < v->nStates - 1) {
    double fW = vit_fweight(frame[i * 2]);
    return (-log(pv) + thr) * v->wLocal + fW * v->wRange;
  }
  double flag = 0.0;
  for (int j = 0; j < v->nStates; j++) if (frame[j * 2 + 1] >= v->voiceThresh) { flag = v->wThr; break; }
  return v->wLocal * flag;
}

/* :224-252; i = state in the current frame, j = state in the previous frame */
static double vit_trans(vit *v, int i, int j, const float *prevF, const float *curF)
{
  const int last = v->nStates - 1;
  if ((i == j) == last) return v->wTuu;            /* the reference's `i == j == nStates-1` */
  if (i < last && j < last) {
    float f0 = prevF[j * 2], f1 = curF[i * 2];
    if (f0 == 0 || f1 == 0) return 999.0;
    double r = log((double)(f1 / f0));
    double x = v->wTvv * fabs(r) + v->wTvvd * fabs(r - v->lastChange);
    v->lastChange = r;
    return x;
  }
  if ((i == last && j < last) || (i < last && j == last)) { v->lastChange = 0.0; return v->wTvuv; }
  return 1.0;
}

static float vit_state_value(const vit *v, int i, const float *frame) { return i < v->nStates - 1 ? frame[i * 2] : 0.0f; }

/* lld/pitchSmootherViterbi.cpp:79-183 */
static long vit_add(vit *v, const float *frame)
{
  int i, j;
  const long bl = v->buflen;
  if (v->wrIdx - v->rdIdx >= bl) return -1;
  float *b = v->buf + (v->wrIdx % bl) * v->frameSize;
  memcpy(b, frame, sizeof(float) * v->frameSize);
  v->wrIdx++;
  float *a = v->prev; v->prev = b;
  if (v->pathIdx == 0 || v->prev == NULL) {
    v->pathIdx = 0; v->convIdx = -1;
    for (i = 0; i < v->nStates; i++) {
      v->pathCosts[i] = vit_local(v, i, frame);
      v->paths[v->pathBuf][i * bl] = i;
    }
  } else {
    int nb = (v->pathBuf + 1) % 2;
    for (i = 0; i < v->nStates; i++) {
      int minState = 0;
      double minCost;
      minCost = v->pathCostsTemp[0] = vit_trans(v, i, 0, a, b) + v->pathCosts[0];
      for (j = 1; j < v->nStates; j++) {
        v->pathCostsTemp[j] = vit_trans(v, i, j, a, b) + v->pathCosts[j];
        if (v->pathCostsTemp[j] < minCost) { minState = j; minCost = v->pathCostsTemp[j]; }
      }
      v->pathCostsNew[i] = minCost + vit_local(v, i, frame);
      memcpy(v->paths[nb] + i * bl, v->paths[v->pathBuf] + minState * bl, bl * sizeof(int));
      v->paths[nb][i * bl + v->pathIdx % bl] = i;
    }
    double *tmp = v->pathCosts; v->pathCosts = v->pathCostsNew; v->pathCostsNew = tmp;
    v->pathBuf = nb;
  }
  v->pathIdx++;
  if (v->pathIdx - v->convIdx > bl) {
    int minState = 0;
    for (i = 1; i < v->nStates; i++) if (v->pathCosts[i] < v->pathCosts[minState]) minState = i;
    v->convIdx++;
    v->bestPath[v->convIdx % bl] = v->paths[v->pathBuf][minState * bl + v->convIdx % bl];
  } else {
    for (long n = v->convIdx + 1; n < v->pathIdx; n++) {
      int x = v->paths[v->pathBuf][n % bl];
      int match = 1;
      for (i = 1; i < v->nStates; i++) if (x != v->paths[v->pathBuf][i * bl + n % bl]) { match = 0; break; }
      if (!match) break;
      v->convIdx++;
      v->bestPath[v->convIdx % bl] = x;
    }
  }
  return v->convIdx + 1 - v->rdIdx;
}

/* include/lld/pitchSmootherViterbi.hpp:105-125 */
static void vit_flush(vit *v)
{
  int minState = 0;
  for (int i = 1; i < v->nStates; i++) if (v->pathCosts[i] < v->pathCosts[minState]) minState = i;
  for (long i = v->convIdx + 1; i < v->pathIdx; i++) {
    v->convIdx++;
    v->bestPath[v->convIdx % v->buflen] = v->paths[v->pathBuf][minState * v->buflen + v->convIdx % v->buflen];
  }
}

typedef struct { const osm_or_viterbi_cfg *vc; int nc; float lastValidf0; float *out; long nOut; int K; } vit_sink;

/* lld/pitchSmootherViterbi.cpp:470-545 */
static void vit_drain(vit *v, vit_sink *s)
{
  const osm_or_viterbi_cfg *vc = s->vc;
  long avail = v->convIdx + 1 - v->rdIdx;
  for (long k = 0; k < avail; k++) {
    int state = v->bestPath[v->rdIdx % v->buflen];
    const float *b = v->buf + (v->rdIdx % v->buflen) * v->frameSize;
    float f0 = vit_state_value(v, state, b);
    v->rdIdx++;
    float *o = s->out + s->nOut * s->K;
    int n = 0;
    if (vc->F0final) o[n++] = f0;
    if (vc->F0finalLog) {
      float fs = 0.0f;
      /* the reference's C++ resolves log(float) to the float overload: 12 * logf(f0 / 27.5f) / logf(2) in float
       * (verified against oracle/_ref: the double form differs in 1/3 of the frames) */
      if (f0 > 29.136) fs = (float)12.0 * logf(f0 / (float)27.5) / logf((float)2.0);
      else if (f0 > 0.0) fs = 1.0f;
      o[n++] = fs;
    }
    if (vc->F0finalEnv || vc->F0finalEnvLog) {
      if (f0 <= 0.0) f0 = s->lastValidf0; else s->lastValidf0 = f0;
      if (vc->F0finalEnv) o[n++] = f0;
      if (vc->F0finalEnvLog) {
        float fs = 0.0f;
        if (f0 > 29.136) fs = (float)12.0 * logf(f0 / (float)27.5) / logf((float)2.0);
        else if (f0 > 0.0) fs = 1.0f;
        o[n++] = fs;
      }
    }
    float vp = state < s->nc ? b[state * 2 + 1] : b[1];
    if (vc->voicingFinalClipped) o[n++] = vp >= v->voiceThresh ? vp : 0.0f;
    if (vc->voicingFinalUnclipped) o[n++] = vp;
    s->nOut++;
  }
}

long osm_or_viterbi(const osm_or_pitchshs_cfg *ps, const osm_or_viterbi_cfg *vc, const float *shs, long T, float *out,
                    long *n_before_eoi)
{
  const int nc = ps->nCandidates < 1 ? 1 : (ps->nCandidates > 20 ? 20 : ps->nCandidates);
  const int Kin = osm_or_pitchshs_num_out(ps);
  if (!ps->voicing) return -1;
  vit v; memset(&v, 0, sizeof v);
  v.nStates = nc + 1; v.buflen = vc->bufferLength; v.frameSize = nc * 2 + 4;      /* :459-461 */
  v.convIdx = -1;
  v.buf = (float *)malloc(sizeof(float) * v.frameSize * v.buflen);
  v.paths[0] = (int *)malloc(sizeof(int) * v.nStates * v.buflen);
  v.paths[1] = (int *)malloc(sizeof(int) * v.nStates * v.buflen);
  v.bestPath = (int *)malloc(sizeof(int) * v.nStates * v.buflen);
  v.pathCosts = (double *)calloc(v.nStates, sizeof(double));
  v.pathCostsNew = (double *)calloc(v.nStates, sizeof(double));
  v.pathCostsTemp = (double *)calloc(v.nStates, sizeof(double));
  v.voiceThresh = (float)ps->voicingCutoff;                                        /* level meta data, pitchBase.cpp:150-153 */
  /* setWeights (hpp:291-299) stores tvv in wTvvd, the wTvvd option is not used */
  v.wLocal = vc->wLocal; v.wTvv = vc->wTvv; v.wTvvd = vc->wTvv; v.wTvuv = vc->wTvuv; v.wThr = vc->wThr;
  v.wRange = vc->wRange; v.wTuu = vc->wTuu; v.lastChange = 1.0;
  vit_sink s; memset(&s, 0, sizeof s);
  s.vc = vc; s.nc = nc; s.out = out; s.K = osm_or_viterbi_num_out(vc);
  float *frame = (float *)calloc(v.frameSize, sizeof(float));
  for (long t = 0; t < T; t++) {                                                   /* :437-455 */
    const float *row = shs + t * Kin;
    for (int i = 0; i < nc; i++) { frame[i * 2] = row[1 + i]; frame[i * 2 + 1] = row[1 + nc + i]; }
    frame[nc * 2] = 0.0f; frame[nc * 2 + 1] = 0.0f; frame[nc * 2 + 2] = 0.0f;
    frame[nc * 2 + 3] = (float)t;
    vit_add(&v, frame);
    vit_drain(&v, &s);
  }
  if (n_before_eoi) *n_before_eoi = s.nOut;      /* frames the level holds when the end of input is signalled */
  vit_flush(&v);                                                                   /* :431-434 */
  vit_drain(&v, &s);
  free(frame); free(v.buf); free(v.paths[0]); free(v.paths[1]); free(v.bestPath);
  free(v.pathCosts); free(v.pathCostsNew); free(v.pathCostsTemp);
  return s.nOut;
}

/* ------------------------------------------------------------------ cValbasedSelector */

void osm_or_valbased_select(const float *sel, const float *x, long T, int K, double threshold, double outputVal, float *out)
{
  const float thr = (float)threshold, ov = (float)outputVal;          /* valbasedSelector.cpp:95,101 */
  for (long t = 0; t < T; t++)
    for (int k = 0; k < K; k++) out[t * K + k] = sel[t] > thr ? x[t * K + k] : ov;   /* :192-232 */
}

/* ------------------------------------------------------------------ cPitchJitter */

int osm_or_jitter_num_out(const osm_or_jitter_cfg *jc)
{
  return (jc->jitterLocal ? 1 : 0) + (jc->jitterDDP ? 1 : 0) + (jc->jitterLocalEnv ? 1 : 0) + (jc->jitterDDPEnv ? 1 : 0) +
         (jc->shimmerLocal ? 1 : 0) + (jc->shimmerLocalDB ? 1 : 0) + (jc->shimmerLocalEnv ? 1 : 0) + (jc->shimmerLocalDBEnv ? 1 : 0) +
         (jc->harmonicERMS ? 1 : 0) + (jc->noiseERMS ? 1 : 0) + (jc->linearHNR ? 1 : 0) + (jc->logHNR ? 1 : 0) +
         (jc->refinedF0 ? 1 : 0) + (jc->sourceQualityMean ? 1 : 0) + (jc->sourceQualityRange ? 1 : 0);
}

/* lld/pitchJitter.cpp:339-413 */
static double cross_corr(const float *x, const float *y, long N)
{
  double cc = 0.0, mx = 0.0, my = 0.0, nx = 0, ny = 0;
  for (long i = 0; i < N; i++) { mx += x[i]; my += y[i]; }
  mx /= (double)N; my /= (double)N;
  for (long i = 0; i < N; i++) {
    cc += (x[i] - mx) * (y[i] - my);
    nx += (x[i] - mx) * (x[i] - mx);
    ny += (y[i] - my) * (y[i] - my);
  }
  cc /= sqrt(nx) * sqrt(ny);
  return cc;
}

/* :418-456 */
static float amp_diff(const float *x, long Nx, const float *y, long Ny, double *maxI0, double *maxI1, float *A0o, float *A1o)
{
  double A0 = 1.0, A1 = 1.0;
  long mI = 1;
  float max0 = x[1], min0 = x[1];
  for (long i = 1; i < Nx - 1; i++) { if (x[i] > max0) { max0 = x[i]; mI = i; } if (x[i] < min0) min0 = x[i]; }
  *maxI0 = quad3((double)(mI - 1), x[mI - 1], (double)mI, x[mI], (double)(mI + 1), x[mI + 1], &A0, NULL);
  mI = 1;
  float max1 = y[1], min1 = y[1];
  for (long i = 1; i < Ny - 1; i++) { if (y[i] > max1) { max1 = y[i]; mI = i; } if (y[i] < min1) min1 = y[i]; }
  *maxI1 = quad3((double)(mI - 1), y[mI - 1], (double)mI, y[mI], (double)(mI + 1), y[mI + 1], &A1, NULL);
  *A0o = max0 - min0; *A1o = max1 - min1;
  return (float)fabs((max0 - min0) - (max1 - min1));
}

/* :461-513 */
static float rms_amp_diff(const float *x, long Nx, const float *y, long Ny, double *maxI0, double *maxI1, float *A0o, float *A1o)
{
  double A0 = 1.0, A1 = 1.0;
  long i, mI = 1;
  float mx = x[1];
  float rmsX = x[0] * x[0];
  for (i = 1; i < Nx - 1; i++) { if (x[i] > mx) { mx = x[i]; mI = i; } rmsX += x[i] * x[i]; }
  rmsX = sqrt((rmsX + x[i] * x[i]) / (float)Nx);
  *maxI0 = quad3((double)(mI - 1), x[mI - 1], (double)mI, x[mI], (double)(mI + 1), x[mI + 1], &A0, NULL);
  mI = 1; mx = y[1];
  float rmsY = y[0] * y[0];
  for (i = 1; i < Ny - 1; i++) { if (y[i] > mx) { mx = y[i]; mI = i; } rmsY += y[i] * y[i]; }
  rmsY = sqrt((rmsY + y[i] * y[i]) / (float)Ny);
  *maxI1 = quad3((double)(mI - 1), y[mI - 1], (double)mI, y[mI], (double)(mI + 1), y[mI + 1], &A1, NULL);
  *A0o = rmsX; *A1o = rmsY;
  return (float)fabs(rmsX - rmsY);
}

static double amp_ratio_db(double a) { return a > 10e-50 ? 20.0 * log(a) / log(10.0) : -1000.0; }   /* smileUtil.c:2066-2073 */

long osm_or_pitch_jitter(const osm_or_frontend *fe, const osm_or_jitter_cfg *jc, const int16_t *pcm, long L,
                         int n_chan, const float *F0in, long T, float *out)
{
  const long Nfr = osm_or_frame_size_samples(fe), H = osm_or_frame_step_samples(fe);
  const int K = osm_or_jitter_num_out(jc);
  if (T <= 0 || K <= 0) return 0;
  float *wav = (float *)malloc(sizeof(float) * (L > 0 ? L : 1));
  osm_or_pcm16_to_float(pcm, L, n_chan, wav);
  const double Ts = 1.0 / fe->sample_rate;          /* wave level period */
  const double pitchT = fe->frame_step_sec;         /* period of the F0 level (winToVecProcessor.cpp:563) */
  const int minNumPeriods = jc->minNumPeriods < 2 ? 2 : jc->minNumPeriods;
  float threshCC = (float)jc->minCC;
  if (threshCC < (float)0.01) threshCC = (float)0.01;
  if (threshCC > (float)0.99) threshCC = (float)0.99;
  const float lgHNRfloor = (float)jc->lgHNRfloor;
  /* state (ctor :88-93) */
  long lastIdx = 0, lastMis = 0;
  float lastT0 = 0.0f, lastDiff = 0.0f, lastJitterLocal = 0.0f, lastJitterDDP = 0.0f, lastShimmerLocal = 0.0f;
  float lastJitterLocal_b = 0.0f, lastJitterDDP_b = 0.0f, lastShimmerLocal_b = 0.0f;
  long nOut = 0;
  for (long t = 0; t < T; t++) {
    float F0 = F0in[t];
    /* time meta of frame t of the framer level: time = first sample * Ts, lengthSec from
     * cMatrix::squashTimeMeta (core/dataMemoryLevel.cpp:617-625), framePeriod = Ts */
    const long s0 = t * H;
    const double time = (double)s0 * Ts;
    const double lengthSec = (double)(s0 + Nfr - 1) * Ts - (double)s0 * Ts + Ts;
    long lenF = (long)ceil(lengthSec / Ts);                               /* :609 */
    long startVidx = (long)round(time / Ts);                              /* :612 */
    long ppLen = (long)ceil(pitchT / Ts);                                 /* :616 */
    long toRead0 = ppLen + lastMis, toRead = toRead0;                     /* :624-625 */
    double T0 = 0.0, Tf = 0.0, T0min = 0.0, T0max = 0.0;
    long T0f = 0, T0minF = 0, T0maxF = 0, two_pp = 0;
    if (F0 > 0.0) {                                                       /* :635-648 */
      T0 = 1.0 / F0;
      Tf = T0 / Ts;
      T0f = (long)round(Tf);
      T0min = (1.0 - jc->searchRangeRel) * Tf;
      T0max = (1.0 + jc->searchRangeRel) * Tf;
      T0minF = (long)floor(T0min);
      T0maxF = (long)ceil(T0max);
      two_pp = minNumPeriods * T0maxF + minNumPeriods;
      if (toRead < two_pp) toRead = two_pp;
    }
    long maxRead = lastMis + lenF;                                        /* :649 */
    if (toRead > maxRead) toRead = maxRead;
    if (startVidx - lastMis != lastIdx) {                                 /* :658-663 */
      lastIdx = startVidx;
      if (toRead > lenF) toRead = lenF;
      if (maxRead > lenF) maxRead = lenF;
    }
    if (lastIdx + toRead > L) {          /* the reference's getMatrix fails here: frame skipped (:668-673) */
      lastIdx += toRead0;
      continue;
    }
    const float *data = wav + lastIdx;
    const long nT = toRead;
    float nPeriodsLocal = 0, nPeriodsDDP = 0, nPeriods = 0, avgPeriod = 0.0f, JitterDDP = 0.0f, JitterLocal = 0.0f;
    float avgAmp = 0.0f, avgAmpDiff = 0.0f, eH = 0.0f, eN = 0.0f, HNR = 0.0f, lgHNR = 0.0f, sumCC = 0.0f, maxCC = -2.0f, minCC = -2.0f;
    long start = 0, lastPeriod = 0, i;
    if (F0 > 0.0) {
      int numPeriods = 0;
      long *periodBuffer = (long *)calloc((size_t)(T0f > 0 ? (maxRead / T0minF + 3) : (maxRead + 2)) + 4, sizeof(long));
      float *avgWf = (float *)calloc((size_t)(T0f + 1), sizeof(float));
      double *cc = (double *)calloc((size_t)(T0maxF - T0minF) + 1, sizeof(double));
      long os = start, pp = 0;
      while (start < nT - 2 * T0maxF - 1) {                               /* :728 */
        for (long tf = T0minF; tf <= T0maxF; tf++) cc[tf - T0minF] = cross_corr(data + start, data + start + tf, tf);
        double mx = cc[T0f - T0minF];
        long maxI = -1;
        for (i = 1; i < T0maxF - T0minF - 1; i++) {                       /* :743-754 */
          if (cc[i - 1] < cc[i] && cc[i] > cc[i + 1]) {
            if (maxI == -1) { maxI = i; mx = cc[i]; }
            else if (cc[i] > mx) { maxI = i; mx = cc[i]; }
          }
        }
        pp = maxI == -1 ? T0f : T0minF + maxI;
        os = start;
        if (maxI >= 0) {
          start += pp;
          double max0 = 0.0, max1 = 0.0;
          float a0 = 0.0f, a1 = 0.0f, ad;
          if (jc->shimmerUseRmsAmplitude) ad = rms_amp_diff(data + os, pp, data + start, pp, &max0, &max1, &a0, &a1);
          else ad = amp_diff(data + os, pp, data + start, pp, &max0, &max1, &a0, &a1);
          periodBuffer[numPeriods++] = os;
          for (i = 0; i < T0f; i++) avgWf[i] += data[os + i];
          double conf = 0.0, ccI = 0.0;
          double maxId = fabs(((double)T0minF + quad3((double)(maxI - 1), cc[maxI - 1], (double)maxI, cc[maxI],
                                                      (double)(maxI + 1), cc[maxI + 1], &ccI, &conf))) * Ts;
          sumCC += (float)ccI;
          if (minCC == (float)-2.0 || minCC > (float)ccI) minCC = (float)ccI;
          if (maxCC == (float)-2.0 || maxCC < (float)ccI) maxCC = (float)ccI;
          if (jc->useBrokenJitterThresh) threshCC = minCC;                /* :811-816 */
          if (ccI > threshCC) {
            float period;
            if (jc->usePeakToPeakPeriodLength) period = (float)(((double)start + max1 - (double)os - max0) * Ts);
            else period = (float)maxId;
            avgPeriod += period;
            nPeriods += 1.0f;
            if (lastT0 > 0.0) {
              float diff = (float)fabs(lastT0 - period);
              JitterLocal += diff;
              nPeriodsLocal += 1.0f;
              if (lastDiff > 0.0) { JitterDDP += fabs(lastDiff - diff); nPeriodsDDP += 1.0f; }
              lastDiff = diff;
            }
            lastT0 = period;
            avgAmp += (a0 + a1) / (float)2.0;
            avgAmpDiff += ad;
          }
        } else {
          start += T0f;
        }
        if (start < toRead0 - 1) lastPeriod = start;                      /* :856-858 */
      }
      periodBuffer[numPeriods++] = start;
      float Eh = 0.0f;
      for (i = 0; i < T0f && start + i < nT; i++) {                       /* :865-870 */
        avgWf[i] += data[start + i];
        avgWf[i] /= (float)numPeriods;
        if (i > 2 && i < T0f - 2) Eh += avgWf[i] * avgWf[i];
      }
      if (T0f - 4 > 0) Eh /= (float)(T0f - 4);
      Eh = sqrt(Eh);
      float En = 0.0f; long nEn = 0;
      if (pp > 0) periodBuffer[numPeriods] = start + pp;
      for (i = 0; i < numPeriods; i++) {                                  /* :882-889 */
        long n = 2;
        long hi = periodBuffer[i + 1] < periodBuffer[i] + T0f ? periodBuffer[i + 1] : periodBuffer[i] + T0f;
        for (long j = periodBuffer[i] + 2; j < hi - 2; j++) {
          float delta = data[j] - avgWf[n++];
          En += delta * delta;
          nEn++;
        }
      }
      if (nEn > 0) En /= (float)nEn;
      En = sqrt(En);
      eH = Eh; eN = En;
      if (En > 0.0) {
        HNR = Eh / En;
        if (HNR > 0.0) lgHNR = (float)(20.0 * log((double)HNR) / log(10.0));
        else lgHNR = lgHNRfloor;
      }
      if (numPeriods > 0) sumCC /= (float)numPeriods;
      lastMis = toRead0 - lastPeriod;
      free(cc); free(periodBuffer); free(avgWf);
    } else {                                                              /* :918-943 */
      start = toRead0; lastPeriod = toRead0; lastMis = 0;
      lastT0 = 0.0f; lastDiff = 0.0f; lastJitterDDP = 0.0f; lastJitterLocal = 0.0f; lastShimmerLocal = 0.0f;
      if (jc->noiseERMS || jc->linearHNR || jc->logHNR) {
        double E = 0.0;
        for (i = 0; i < nT; i++) E += data[i] * data[i];
        E /= (double)nT;
        eH = 0.0f; HNR = 0.0f; eN = (float)sqrt(E); lgHNR = lgHNRfloor;
      }
    }
    lastIdx += lastPeriod;
    if (jc->onlyVoiced && F0 == 0.0) continue;
    float *o = out + nOut * K;
    int n = 0;
    const int okL = nPeriods > 0.0 && nPeriodsLocal > 0.0 && F0 > 0.0;
    if (okL) { JitterLocal /= nPeriodsLocal; lastJitterLocal_b = lastJitterLocal = JitterLocal / (avgPeriod / nPeriods); }
    if (jc->jitterLocal) {
      if (okL) { if (lastJitterLocal > 1.0) lastJitterLocal = 1.0f; o[n] = lastJitterLocal; }
      else if (nPeriods == 0.0 && F0 > 0.0) { if (lastJitterLocal > 1.0) lastJitterLocal = 1.0f; o[n] = lastJitterLocal; }
      else o[n] = 0.0f;
      n++;
    }
    if (jc->jitterLocalEnv) { if (lastJitterLocal_b > 1.0) lastJitterLocal_b = 1.0f; o[n++] = lastJitterLocal_b; }
    const int okD = nPeriods > 0.0 && nPeriodsDDP > 0.0 && F0 > 0.0;
    if (okD) { JitterDDP /= nPeriodsDDP; lastJitterDDP_b = lastJitterDDP = JitterDDP / (avgPeriod / nPeriods); }
    if (jc->jitterDDP) {
      if (okD) { if (lastJitterDDP > 1.0) lastJitterDDP = 1.0f; o[n] = lastJitterDDP; }
      else if (nPeriods == 0.0 && F0 > 0.0) { if (lastJitterDDP > 1.0) lastJitterDDP = 1.0f; o[n] = lastJitterDDP; }
      else o[n] = 0.0f;
      n++;
    }
    if (jc->jitterDDPEnv) { if (lastJitterDDP_b > 1.0) lastJitterDDP_b = 1.0f; o[n++] = lastJitterDDP_b; }
    if (nPeriods > 0.0 && F0 > 0.0) {
      if (avgAmp > 0.0) lastShimmerLocal_b = lastShimmerLocal = (avgAmpDiff / avgAmp);
      else lastShimmerLocal = 0.0f;
    }
    if (jc->shimmerLocal || jc->shimmerLocalDB) {
      if (F0 > 0.0) {          /* nPeriods > 0 and nPeriods == 0 branches are identical (:1009-1031) */
        if (lastShimmerLocal > 1.0) lastShimmerLocal = 1.0f;
        if (jc->shimmerLocal) o[n++] = lastShimmerLocal;
        if (jc->shimmerLocalDB) o[n++] = (float)amp_ratio_db(lastShimmerLocal + 1.0);
      } else {
        if (jc->shimmerLocal) o[n++] = 0.0f;
        if (jc->shimmerLocalDB) o[n++] = 0.0f;
      }
    }
    if (jc->shimmerLocalEnv) { if (lastShimmerLocal_b > 1.0) lastShimmerLocal_b = 1.0f; o[n++] = lastShimmerLocal_b; }
    if (jc->harmonicERMS) o[n++] = eH;
    if (jc->noiseERMS) o[n++] = eN;
    if (jc->linearHNR) o[n++] = HNR;
    if (jc->logHNR) { if (lgHNR < lgHNRfloor) lgHNR = lgHNRfloor; o[n++] = lgHNR; }
    if (jc->refinedF0) o[n++] = (nPeriods > 0.0 && F0 > 0.0) ? (float)1.0 / (avgPeriod / nPeriods) : 0.0f;
    if (jc->sourceQualityMean) o[n++] = sumCC;
    if (jc->sourceQualityRange) o[n++] = fabs(maxCC - minCC);
    nOut++;
  }
  free(wav);
  return nOut;
}
