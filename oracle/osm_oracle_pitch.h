/*
 * osm_oracle_pitch.h -- CPU restatement of the sub-harmonic-summation pitch chain of the
 * ComParE_2016 / eGeMAPS graphs:
 *   cSpecScale -> cPitchShs -> cPitchSmootherViterbi -> cValbasedSelector -> cPitchJitter,
 * plus cDeltaRegression with onlyInSegments=1 (the deltas of those columns).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as osm_oracle.h): only tests/, smoke() and bench.py's
 * cpu_baseline leg may load it.  Pinned against the UNMODIFIED reference (oracle/_ref) through
 * level taps (scripts/make_golden_pitch.py -> tests/golden/pitch_goldens.npz).
 * Citations are relative to /root/reference/src.
 */
#ifndef OSM_ORACLE_PITCH_H
#define OSM_ORACLE_PITCH_H
#include <stdint.h>
#include "osm_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {             /* cSpecScale (dsp/specScale.cpp), scale=octave, sourceScale=lin, spline */
  double minF, maxF;         /* 25, -1 */
  int nPointsTarget;         /* 0 -> number of magnitude bins */
  int specSmooth, specEnhance, auditoryWeighting;
} osm_or_specscale_cfg;

typedef struct {             /* cPitchShs on cPitchBase (lld/pitchShs.cpp, lldcore/pitchBase.cpp) */
  double maxPitch, minPitch; /* 620, 52 */
  int nCandidates;           /* 3 (ComParE: 6) */
  int scores, voicing, F0C1, voicingC1, F0raw, voicingClip;
  double voicingCutoff;      /* 0.70 */
  int octaveCorrection;
  int nHarmonics;            /* 15 */
  double compressionFactor;  /* 0.85 */
  int greedyPeakAlgo;
  double lfCut;
} osm_or_pitchshs_cfg;

typedef struct {             /* cPitchSmootherViterbi (lld/pitchSmootherViterbi.cpp) */
  int bufferLength;          /* 30 */
  int F0final, F0finalLog, F0finalEnv, F0finalEnvLog, voicingFinalClipped, voicingFinalUnclipped;
  double wLocal, wTvv, wTvvd, wTvuv, wThr, wRange, wTuu;
} osm_or_viterbi_cfg;

typedef struct {             /* cPitchJitter (lld/pitchJitter.cpp) */
  double searchRangeRel;     /* 0.10 (ComParE: 0.25) */
  int jitterLocal, jitterDDP, jitterLocalEnv, jitterDDPEnv;
  int shimmerLocal, shimmerLocalDB, shimmerLocalEnv, shimmerLocalDBEnv;
  int harmonicERMS, noiseERMS, linearHNR, logHNR;
  double lgHNRfloor;         /* -100 */
  int shimmerUseRmsAmplitude;
  int minNumPeriods;         /* 2 */
  double minCC;              /* 0.5 */
  int refinedF0, sourceQualityRange, sourceQualityMean;
  int usePeakToPeakPeriodLength, useBrokenJitterThresh, onlyVoiced;
} osm_or_jitter_cfg;

int osm_or_pitchshs_num_out(const osm_or_pitchshs_cfg *ps);
int osm_or_viterbi_num_out(const osm_or_viterbi_cfg *vc);
int osm_or_jitter_num_out(const osm_or_jitter_cfg *jc);

/* cSpecScale + cPitchShs over all frames of one utterance.  out_shs = [T][pitchshs_num_out],
 * tap_hps (optional) = [T][nPoints] scaled spectrum.  Returns T. */
long osm_or_pitch_shs(const osm_or_frontend *fe, const osm_or_specscale_cfg *sc, const osm_or_pitchshs_cfg *ps,
                      const int16_t *pcm, long n_samples, int n_chan, float *out_shs, float *tap_hps);

/* cPitchSmootherViterbi over the cPitchShs level of one utterance: out = [T][viterbi_num_out];
 * *n_before_eoi (optional) = frames written before the trellis is flushed at the end of input (the
 * reference's downstream levels see the remaining ones only during / after its first EOI pass) */
long osm_or_viterbi(const osm_or_pitchshs_cfg *ps, const osm_or_viterbi_cfg *vc, const float *shs, long T, float *out,
                    long *n_before_eoi);

/* cValbasedSelector (other/valbasedSelector.cpp:153-233) with idx=0, removeIdx=1, zeroVec=1:
 * rows whose selector value is not > threshold become outputVal */
void osm_or_valbased_select(const float *sel, const float *x, long T, int K, double threshold, double outputVal, float *out);

/* cPitchJitter over one utterance: F0 = [T] (one value per frame of `fe`), out = [T][jitter_num_out];
 * returns the number of frames written */
long osm_or_pitch_jitter(const osm_or_frontend *fe, const osm_or_jitter_cfg *jc, const int16_t *pcm, long n_samples,
                         int n_chan, const float *F0, long T, float *out);

/* cDeltaRegression with onlyInSegments=1 on the whole level [T][K] (input frames before EOI = n0):
 * the reference's norm accumulates over every frame and element ever processed
 * (dspcore/deltaRegression.cpp:123-141).  out = [T+win][K] */
long osm_or_delta_segments(const float *in, long T, long n0, int K, int win, float *out);

#ifdef __cplusplus
}
#endif
#endif
