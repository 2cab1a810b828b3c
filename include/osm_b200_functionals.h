/*
 * osm_b200_functionals.h -- cFunctionals on the GPU: per-utterance summaries of resident LLD rows.
 *
 * Mirrors the reference's cFunctionals component in full-input mode (frameMode = full: one output vector per utterance,
 * src/functionals/functionals.cpp:284-330) with the sub-components
 *     cFunctionalExtremes     src/functionals/functionalExtremes.cpp:89-132
 *     cFunctionalMeans        src/functionals/functionalMeans.cpp:104-262
 *     cFunctionalMoments      src/functionals/functionalMoments.cpp:89-168
 *     cFunctionalPercentiles  src/functionals/functionalPercentiles.cpp:299-430
 *     cFunctionalRegression   src/functionals/functionalRegression.cpp:141-428
 *     cFunctionalTimes        src/functionals/functionalTimes.cpp:245-371
 *     cFunctionalLpc          src/functionals/functionalLpc.cpp:98-125
 *     cFunctionalSegments     src/functionals/functionalSegments.cpp:305-367 (relTh), :658-797 (nonX / eqX), :800-960
 *     cFunctionalPeaks2       src/functionals/functionalPeaks2.cpp:296-915
 * Field names and defaults are the reference's configuration fields (`[x:cFunctionals]` section: functionalsEnabled,
 * nonZeroFuncts, functNameAppend, masterTimeNorm, and `<Functional>.<field>` for the sub-components).
 *
 * The input is the row-major LLD matrix a plan leaves in HBM ([sum rows][row_stride] float32, utterance u owns rows
 * row_offsets[u] .. row_offsets[u] + n_rows[u]); the output is one row of num_elements floats per utterance, laid out like the
 * reference's functionals level: for every input element, its enabled values in functionalsEnabled order
 * (src/functionals/functionals.cpp:215-256 for the names).  One warp per (utterance, element).
 * No CPU fallback: device < 0 gives a description-only object (names / counts) that cannot run.
 */
#ifndef OSM_B200_FUNCTIONALS_H
#define OSM_B200_FUNCTIONALS_H

#include "osm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  OSM_B200_F_EXTREMES = 0, OSM_B200_F_MEANS, OSM_B200_F_MOMENTS, OSM_B200_F_PERCENTILES, OSM_B200_F_REGRESSION,
  OSM_B200_F_TIMES, OSM_B200_F_LPC, OSM_B200_F_SEGMENTS, OSM_B200_F_PEAKS2,
  OSM_B200_F_ONSET, OSM_B200_F_PEAKS, OSM_B200_F_CROSSINGS, OSM_B200_F_SAMPLES, OSM_B200_F_DCT,
  OSM_B200_F_COUNT_
} osm_b200_functional_type;

/* time normalisation (src/include/functionals/functionalComponent.hpp:27-35) */
#define OSM_B200_TIMENORM_UNSET   (-1)
#define OSM_B200_TIMENORM_SEGMENT 0
#define OSM_B200_TIMENORM_SECOND  1
#define OSM_B200_TIMENORM_FRAME   2

#define OSM_B200_F_MAX_ENABLED 8
#define OSM_B200_F_MAX_PCTL 8
#define OSM_B200_F_MAX_THRESH 8
#define OSM_B200_F_MAX_LPC 16
#define OSM_B200_F_PEAKS2_VALUES 32
#define OSM_B200_F_MAX_SAMPLES 16
#define OSM_B200_F_MAX_DCT 32

/* cFunctionalSegments.segmentationAlgorithm (the three the shipped ComParE_2016 / GeMAPS blocks use) */
#define OSM_B200_SEG_RELTH 0
#define OSM_B200_SEG_NONX  1
#define OSM_B200_SEG_EQX   2
#define OSM_B200_SEG_NARELTH 3   /* relTh on the samples themselves, no 3-frame running average (process_SegThreshNoavg, functionalSegments.cpp:369-413) */

typedef struct {
  /* [x:cFunctionals] */
  int32_t n_enabled;                               /* entries of functionalsEnabled */
  int32_t enabled[OSM_B200_F_MAX_ENABLED];         /* osm_b200_functional_type, in the order of the array */
  int32_t nonZeroFuncts;                           /* 0; 1 = values != 0 only; 2 = values > 0 only */
  int32_t masterTimeNorm;                          /* OSM_B200_TIMENORM_*, UNSET when the field is absent */
  char    functNameAppend[OSM_B200_NAME_LEN];      /* "" = none */
  struct {                                         /* Extremes.* : 1,1,1,1,1,0,1,1 ; norm "frames" */
    int32_t max, min, range, maxpos, minpos, amean, maxameandist, minameandist;
    int32_t norm, normIsSet;                       /* own `norm` field and whether the configuration sets it */
  } extremes;
  struct {                                         /* Means.* : 1,1,1,1,1,1,1,1,0,... ; norm "frames" */
    int32_t amean, absmean, qmean, nzamean, nzabsmean, nzqmean, nzgmean, nnz, flatness, posamean, negamean,
            posqmean, posrqmean, negqmean, negrqmean, rqmean, nzrqmean;
    int32_t norm, normIsSet;
  } means;
  struct {                                         /* Moments.* : 1,1,1,1,0, stddevNorm 0 (1 = / |mean|, 2 = / mean) */
    int32_t variance, stddev, skewness, kurtosis, amean, stddevNorm, doRatioLimit;
  } moments;
  struct {                                         /* Percentiles.* */
    int32_t quartile1, quartile2, quartile3, iqr12, iqr23, iqr13;
    int32_t n_percentile;  double percentile[OSM_B200_F_MAX_PCTL];
    int32_t n_pctlrange;   int32_t pctlrange[OSM_B200_F_MAX_PCTL][2];
    int32_t interp;                                /* 1 */
  } percentiles;
  struct {                                         /* Regression.* : nine values + centroid on by default */
    int32_t linregc1, linregc2, linregerrA, linregerrQ, qregc1, qregc2, qregc3, qregerrA, qregerrQ, centroid;
    int32_t centroidNorm;                          /* SEGMENT */
    int32_t centroidUseAbsValues, centroidRatioLimit;   /* 1, 1 */
    int32_t normRegCoeff, normInputs, oldBuggyQerr, doRatioLimit;   /* 0, 0, 1, 0 */
  } regression;
  struct {                                         /* Times.* : every value on; norm "segment"; buggySecNorm 1 */
    int32_t upleveltime25, downleveltime25, upleveltime50, downleveltime50, upleveltime75, downleveltime75, upleveltime90,
            downleveltime90, risetime, falltime, leftctime, rightctime, duration;
    int32_t buggySecNorm;
    int32_t norm, normIsSet;
  } times;
  struct {                                         /* Lpc.* : lpGain 0, lpc 1, firstCoeff 0, order 5 */
    int32_t lpGain, lpc, firstCoeff, order;
  } lpc;
  struct {                                         /* Segments.* : every value off; maxNumSeg 20; segMinLng 3; pauseMinLng 2; norm "segment" */
    int32_t numSegments, meanSegLen, maxSegLen, minSegLen, segLenStddev;
    int32_t algorithm;                             /* OSM_B200_SEG_* */
    int32_t maxNumSeg;
    int32_t n_thresholds; float thresholds[OSM_B200_F_MAX_THRESH];   /* relTh: relative to the contour's range */
    float   X; int32_t XisRel;                     /* nonX / eqX */
    int32_t segMinLng, segMinLngIsSet, pauseMinLng;
    int32_t norm, normIsSet;
  } segments;
  struct {                                         /* Peaks2.* : every value off; norm "frames"; relThresh 0.1; doRatioLimit 1 */
    int32_t value[OSM_B200_F_PEAKS2_VALUES];       /* in the reference's output order: numPeaks, meanPeakDist, meanPeakDistDelta,
                                                      peakDistStddev, peakRangeAbs, peakRangeRel, peakMeanAbs, peakMeanMeanDist, peakMeanRel,
                                                      ptpAmpMeanAbs, ptpAmpMeanRel, ptpAmpStddevAbs, ptpAmpStddevRel, minRangeAbs, minRangeRel,
                                                      minMeanAbs, minMeanMeanDist, minMeanRel, mtmAmpMeanAbs, mtmAmpMeanRel, mtmAmpStddevAbs,
                                                      mtmAmpStddevRel, meanRisingSlope, maxRisingSlope, minRisingSlope, stddevRisingSlope,
                                                      meanFallingSlope, maxFallingSlope, minFallingSlope, stddevFallingSlope, covFallingSlope,
                                                      covRisingSlope (functionalPeaks2.cpp:24-73) */
    float   relThresh, absThresh; int32_t useAbsThresh, dynRelThresh, doRatioLimit;
    int32_t norm, normIsSet;
  } peaks2;
  struct {                                         /* Onset.* (functionalOnset.cpp:43-54): 0,0,1,0,0 ; thresholds 0 ; norm "segment" */
    int32_t onsetPos, offsetPos, numOnsets, numOffsets, onsetRate;
    float   thresholdOnset, thresholdOffset;       /* `threshold` sets both, thresholdOnset / thresholdOffset override it */
    int32_t useAbsVal;
    int32_t norm, normIsSet;
  } onset;
  struct {                                         /* Peaks.* (functionalPeaks.cpp:45-53): 1,1,1,1,0 ; norm "frames"; overlapFlag = 1 only */
    int32_t numPeaks, meanPeakDist, peakMean, peakMeanMeanDist, peakDistStddev;
    int32_t norm, normIsSet;
  } peaks;
  struct {                                         /* Crossings.* (functionalCrossings.cpp:42-46): 1,1,0 */
    int32_t zcr, mcr, amean;
  } crossings;
  struct {                                         /* Samples.samplepos[] (functionalSamples.cpp:38-78): relative positions in [0, 1];
                                                      none given = 0, 0.25, 0.5, 0.75, 1 */
    int32_t n_samplepos; double samplepos[OSM_B200_F_MAX_SAMPLES];
  } samples;
  struct {                                         /* DCT.firstCoeff / lastCoeff (functionalDCT.cpp:38-72): 1, 6; nCoeffs overrides lastCoeff */
    int32_t firstCoeff, lastCoeff;
  } dct;
} osm_b200_functionals_spec;

typedef struct osm_b200_functionals osm_b200_functionals;

/* sizeof(osm_b200_functionals_spec) as compiled (bindings check their mirror against it) */
OSM_B200_API int32_t osm_b200_functionals_sizeof_spec(void);
/* the reference's defaults (no functional enabled) */
OSM_B200_API void osm_b200_functionals_defaults(osm_b200_functionals_spec *spec);

/* in_names: the n_in element names of the input level (osm_b200_plan_element_name); input_period: frame period of that
 * level in seconds (osm_b200_plan_frame_period).  device < 0: description only. */
OSM_B200_API osm_b200_status osm_b200_functionals_create(const osm_b200_functionals_spec *spec, int32_t n_in,
                                                         const char *const *in_names, double input_period, int32_t device,
                                                         osm_b200_functionals **f);
OSM_B200_API void            osm_b200_functionals_destroy(osm_b200_functionals *f);
OSM_B200_API int32_t         osm_b200_functionals_num_values(const osm_b200_functionals *f);     /* per input element */
OSM_B200_API int32_t         osm_b200_functionals_num_elements(const osm_b200_functionals *f);   /* n_in * num_values */
OSM_B200_API const char     *osm_b200_functionals_element_name(const osm_b200_functionals *f, int32_t idx);

/* d_rows: device LLD matrix, row_stride floats per row, the first n_in columns are summarised.  row_offsets / n_rows: HOST arrays
 * of n_utt entries (first row and number of rows of every utterance; n_rows[u] = 0 gives an all-zero output row).
 * d_out: device [n_utt][num_elements].  Asynchronous on `stream` (cudaStream_t). */
OSM_B200_API osm_b200_status osm_b200_functionals_run_device(osm_b200_functionals *f, const float *d_rows, int32_t row_stride,
                                                             const int64_t *row_offsets, const int64_t *n_rows, int32_t n_utt,
                                                             float *d_out, void *stream);
/* the same for an input level that is a subset / permutation of the row's columns (several cFunctionals instances on one resident
 * LLD matrix, a cVectorConcat of their outputs behind them): cols = HOST array of n_in column indices (NULL: 0 .. n_in-1);
 * utterance u writes its num_elements values at d_out + u * out_stride (out_stride >= num_elements: the caller lays the instances
 * of a concatenated summary row side by side by offsetting d_out) */
OSM_B200_API osm_b200_status osm_b200_functionals_run_device_cols(osm_b200_functionals *f, const float *d_rows, int32_t row_stride,
                                                                  const int32_t *cols, const int64_t *row_offsets, const int64_t *n_rows,
                                                                  int32_t n_utt, float *d_out, int64_t out_stride, void *stream);
/* The glue the shipped summary graphs put behind their cFunctionals instances (config/gemaps/v01b/GeMAPSv01b_core.func.conf.inc:
 * 133-139 cDataSelector picks and renames summary values, config/egemaps/v02/eGeMAPSv02_core.func.conf.inc:34-42 cVectorOperation
 * turns the mean energy into dB, eGeMAPSv02.conf:31-34 cVectorConcat orders the row): out[r][k] = op_k(in[r][src[k]]) for r < n_rows,
 * k < n_out.  src / op / log_floor: HOST arrays of n_out entries (n_out <= OSM_B200_SUMMARY_MAX_OUT); op = OSM_B200_VOP_*:
 * COPY, DBP = 10 / ln 10 * ln(max(x, log_floor)), DBV = 20 / ln 10 * ln(...) (other/vectorOperation.cpp:508-527, float arithmetic).
 * d_in: device [n_rows][in_stride], d_out: device [n_rows][out_stride].  Asynchronous on `stream`. */
enum { OSM_B200_VOP_COPY = 0, OSM_B200_VOP_DBP = 1, OSM_B200_VOP_DBV = 2 };
#define OSM_B200_SUMMARY_MAX_OUT 320
OSM_B200_API osm_b200_status osm_b200_summary_assemble_device(const float *d_in, int64_t in_stride, const int32_t *src, const int32_t *op,
                                                              const float *log_floor, int32_t n_out, int64_t n_rows, float *d_out,
                                                              int64_t out_stride, void *stream);
/* same with host buffers (copies in, runs, copies out, synchronises) */
OSM_B200_API osm_b200_status osm_b200_functionals_run_host(osm_b200_functionals *f, const float *rows, int32_t row_stride,
                                                           const int64_t *row_offsets, const int64_t *n_rows, int32_t n_utt,
                                                           int64_t total_rows, float *out);

#ifdef __cplusplus
}
#endif
#endif
