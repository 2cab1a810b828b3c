// plan.hpp -- internal representation of a compiled LLD plan (host side) and the parameter
// blocks handed to the CUDA kernels.  Not part of the public ABI (include/osm_b200.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/osm_b200.h"

namespace osm {

// ---------------------------------------------------------------------------------------
// Host-side tables.  Each builder restates the reference's table construction with the same
// float/double casting order (citations in tables.cpp).
// ---------------------------------------------------------------------------------------
struct MelBank {
  int nBands = 0;
  int nBins = 0;
  int nLo = 0, nHi = 0;            // bins [nLo, nHi) are visited (melspec.cpp:543)
  std::vector<float> coef;         // per bin weight of the lower band (melspec.cpp:441-447)
  std::vector<int> chanMap;        // per bin lower band index, -1, -3 (melspec.cpp:427-438)
  std::vector<double> bandHz;      // band centres in Hz (field info, melspec.cpp:408-411)
  // derived for the kernel: range r (0..nBands) = bins whose chanMap == r-1, as [begin,end)
  std::vector<int> rangeBegin;     // size nBands+2 ; rangeBegin[r+1] = end of range r
  float outScale = 1.f;            // htkcompatible scaling (melspec.cpp:559-569)
  bool usePower = false;
};

struct FrontEnd {
  double sampleRate = 0;
  int nChan = 1;
  int format = OSM_B200_PCM_S16;
  bool mixdown = true;
  int frameSize = 0, frameStep = 0, nfft = 0, nBins = 0;
  double frameSizeSec = 0;         // cFramer.frameSize (nominal)
  double frameStepSec = 0;
  double fftFrameSizeSec = 0;      // after cTransformFFT's rescale (transformFft.cpp:78-85)
  bool preemph = false;
  float preK = 0.f;
  int preDe = 0;
  std::vector<float> window;       // (float)win[n], size frameSize (windower.cpp:226)
  float winOffset = 0.f;
  bool zeroPadSymmetric = false;   // phase only; magnitude consumers are unaffected
};

enum StaticOpKind { SOP_MFCC = 0, SOP_PLP, SOP_MELSPEC, SOP_SPECTRAL, SOP_ENERGY, SOP_MZCR };

struct MfccOp {
  int melIdx = 0;
  int first = 0, last = 0, nMfcc = 0;
  float melfloor = 0.f;            // value compared against
  float logMelfloor = 0.f;         // logf(melfloor)
  bool doLog = true;
  std::vector<float> cosT;         // [nMfcc][nBands] rows in OUTPUT order (htk c0-last swap applied)
  std::vector<float> liftFactor;   // [nMfcc] = sintable[i0] * factor, output order
};

struct PlpOp {
  int melIdx = 0;
  int lpOrder = 0, nAuto = 0, nFreq = 0, nCeps = 0, firstCC = 0, lastCC = 0;
  bool doLog = false, doAud = true, doInvLog = false, doIDFT = true, doLP = true, doLpToCeps = true;
  bool htk = true;
  float melfloor = 1.f, logMelfloor = 0.f, compression = 0.33f;
  bool lifter = false;
  std::vector<float> eql;          // [nBands] equal loudness weights (log of them when doLog)
  std::vector<float> cosT;         // [nAuto][nFreq] IDFT table (plp.cpp:298-306)
  std::vector<float> lift;         // [nCeps] lifter per OUTPUT slot (plp.cpp:560-573)
  int nOut = 0;
};

struct StaticOp {
  StaticOpKind kind;
  int outCol = 0, nOut = 0;
  int arrNameOffset = 0;
  MfccOp mfcc;
  PlpOp plp;
};

// temporal stage applied to a static column range (cWindowProcessor family)
enum StageKind { ST_DELTA = 0, ST_SMA = 1 };
struct Stage { StageKind kind; int win; int flags; };

// one contiguous block of output columns
struct OutGroup {
  int srcCol = 0, n = 0;           // columns of the static vector
  std::vector<Stage> stages;       // applied in order
  int outCol = 0;
};

struct PlanDesc {
  FrontEnd fe;
  std::vector<MelBank> mels;
  std::vector<StaticOp> ops;
  int nStatic = 0;
  std::vector<OutGroup> groups;
  int nOut = 0;
  std::vector<std::string> names;  // output element names
};

// graph.cpp
osm_b200_status compile_graph(const osm_b200_component *comps, int n, const char *outputLevel,
                              PlanDesc &out, std::string &err);
int64_t desc_num_frames(const PlanDesc &d, int64_t nSampleFrames);
int64_t desc_num_static_frames(const PlanDesc &d, int64_t nSampleFrames);

// tables.cpp
void build_window(int winFunc, int N, double sigma, double gain, std::vector<float> &w);
void build_mel(const osm_b200_melspec &cfg, int nBins, double frameSizeSec, MelBank &mb);
void build_mfcc(const osm_b200_mfcc &cfg, int nBands, MfccOp &op);
bool build_plp(const osm_b200_plp &cfg, const MelBank &mb, PlpOp &op, std::string &err);

}  // namespace osm
