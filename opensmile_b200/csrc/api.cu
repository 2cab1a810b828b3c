// api.cu -- the C ABI of libosm_b200.so (include/osm_b200.h): plan objects, device tables,
// batch bookkeeping (tile lists, row offsets) and kernel launches.  No CPU fallback: every
// compute entry point fails with OSM_B200_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"

using namespace osm;

// the _f32 builds of the PCM-reading launchers (kernels.cuh: OSM_F32_VARIANT), for plans whose input is not 16-bit integer
namespace osm {
cudaError_t launch_lld_f32(const LldParams &p, int nfft, int numSMs, cudaStream_t st, LldLaunchInfo *info);
cudaError_t launch_energy_f32(const TimeOpParams &p, cudaStream_t st);
cudaError_t launch_mzcr_f32(const TimeOpParams &p, cudaStream_t st);
cudaError_t launch_intensity_f32(const TimeOpParams &p, cudaStream_t st);
cudaError_t launch_formant_f32(const FormantParams &p, cudaStream_t st);
cudaError_t launch_jitter_f32(const JitterParams &p, int u0, int u1, cudaStream_t st);
}

namespace {
thread_local std::string g_err;
}
// shared with the other translation units of the library (functionals.cu): the message osm_b200_last_error() returns
namespace osm { osm_b200_status set_last_error(osm_b200_status st, const std::string &msg) { g_err = msg; return st; } }

namespace {

osm_b200_status fail(osm_b200_status st, const std::string &msg) { return osm::set_last_error(st, msg); }

osm_b200_status cuda_fail(cudaError_t e, const char *what)
{
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error in %s: %s", what, cudaGetErrorString(e));
  return fail(OSM_B200_ERR_CUDA, buf);
}

#define CU(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);       \
  } while (0)

// ---- input formats (cWaveSource.format) ----
int sample_bytes(int format)
{
  switch (format) {
    case OSM_B200_PCM_S8: return 1;
    case OSM_B200_PCM_S16: return 2;
    case OSM_B200_PCM_S24: return 3;
    default: return 4;                 // F32, S24_32, S32
  }
}
// channel count the kernels see: 16-bit input is read in place; every other format is pre-converted to mono floats, whose 4-byte
// sample frame the kernels address as "two int16 per frame" (LldParams::pcmF32)
int kernel_nchan(const FrontEnd &fe) { return fe.format == OSM_B200_PCM_S16 ? fe.nChan : 2; }

// smilePcm_convertSamples / smilePcm_convertFloatSamples with monoMixdown (smileutil/smileUtil.c:2518-2580, 2651-2661): the channel
// values are converted to float and summed in channel order starting from 0.0f, the sum is divided by the channel count, then by
// the format's full scale -- two IEEE divisions, exactly the reference's statement.  One thread per sample frame.
__global__ void __launch_bounds__(256) pcm_convert_kernel(const unsigned char *in, int format, int nChan, long long nFrames, float *out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFrames) return;
  float tmp = 0.0f;
  float scale = 1.0f;
  switch (format) {
    case OSM_B200_PCM_S8: {
      const signed char *b = reinterpret_cast<const signed char *>(in) + i * nChan;
      for (int c = 0; c < nChan; c++) tmp = __fadd_rn(tmp, (float)b[c]);
      scale = 127.0f;
      break;
    }
    case OSM_B200_PCM_S24: {
      const unsigned char *b = in + i * nChan * 3;
      for (int c = 0; c < nChan; c++) {                                             // :2543-2552 (byte assembly, arithmetic shift)
        const unsigned int is = ((unsigned int)b[3 * c] << 8) | ((unsigned int)b[3 * c + 1] << 16) | ((unsigned int)b[3 * c + 2] << 24);
        tmp = __fadd_rn(tmp, (float)((int)is >> 8));
      }
      scale = 8388352.0f;                                                           // (float)(32767.0 * 256.0)
      break;
    }
    case OSM_B200_PCM_S24_32: {
      const int *b = reinterpret_cast<const int *>(in) + i * nChan;
      for (int c = 0; c < nChan; c++) tmp = __fadd_rn(tmp, (float)(b[c] & 0xFFFFFF));   // :2559 (no sign extension)
      scale = 8388352.0f;
      break;
    }
    case OSM_B200_PCM_S32: {
      const int *b = reinterpret_cast<const int *>(in) + i * nChan;
      for (int c = 0; c < nChan; c++) tmp = __fadd_rn(tmp, (float)b[c]);
      scale = 2147483648.0f;                                                        // (float)2147483647.0
      break;
    }
    default: {                                                                      // OSM_B200_PCM_F32: no full-scale division (:2658)
      const float *b = reinterpret_cast<const float *>(in) + i * nChan;
      for (int c = 0; c < nChan; c++) tmp = __fadd_rn(tmp, b[c]);
      out[i] = __fdiv_rn(tmp, (float)nChan);
      return;
    }
  }
  out[i] = __fdiv_rn(__fdiv_rn(tmp, (float)nChan), scale);
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;   // elements
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace


// per-stream runtime state (one framer / FFT chain)
// one lld_kernel launch of a stream: the FFT front end + one band op (or none: magnitude dump only)
struct PassRt {
  LldParams kp;
  unsigned char *dConst = nullptr;
  int op = -1;                   // index into PlanDesc::ops, -1 = no band op
  bool rasta = false;            // cPlp with RASTA: kp stops at the band level, tail = the rest of cPlp
  LldParams tail;
  RastaParams rp;
  DevBuf<float> dBand;           // [static rows][nBands]
};

struct StreamRt {
  unsigned char *dConst = nullptr;
  LldParams kp;                  // pass 0 (the only pass of fused plans)
  std::vector<PassRt> extra;     // passes 1.. (further band ops on the same FFT chain)
  PassRt pass0x;                 // RASTA state of pass 0 (kp / dConst of pass 0 live above)
  int tileF = 32;
  bool runLld = false;           // lld_kernel is launched for this stream
  bool forceNarrow = false;
  bool needTiles = false;        // standalone ops read this stream tile by tile
  const float *dWindow = nullptr;   // [frameSize] window floats (time-domain ops)
  std::vector<int32_t> uttChunk0, uttTile0;
  PinBuf<ChunkRef> hChunks; DevBuf<ChunkRef> dChunks; size_t nChunks = 0;
  PinBuf<OpTile> hTiles; DevBuf<OpTile> dTiles; size_t nTiles = 0;
  DevBuf<float> dMag;
};

struct OpRt {
  int kind = 0, stream = 0;
  int vSrcCol = 0, vN = 0, vOutCol = 0;     // SOP_VECOP
  int magMode = 0; float magN = 1.f, magDbNorm = 0.f, magMinDb = 0.f;   // SOP_MAG: cFFTmagphase normalise / power / dBpsd
  SpectralParams sp;
  TimeOpParams tp;
  AcfPitchParams ap;
  double *dSharpW = nullptr;
  float2 *dTw = nullptr;
  DevBuf<PitchRaw> dRaw;
  // SHS pitch chain (SOP_PITCH) / cPitchJitter (SOP_JITTER)
  ShsParams shs;
  ViterbiParams vit;
  JitterParams jit;
  HarmonicsParams hrm;                  // SOP_HARMONICS
  double *dCosTab = nullptr;            // its cos(2 pi m / N) table
  FormantParams fmt;                    // SOP_FORMANT
  float *dFmtD = nullptr;               // its resampling table
  unsigned char *dPitchTab = nullptr;   // spline / interpolation / harmonic tables of the chain
  DevBuf<float> dShs;                   // [static rows][nShsCols] cPitchShs level
  DevBuf<int> dLag;                     // [nUtt] frames of the Viterbi level before the end-of-input flush
  int descOp = -1;                      // index into PlanDesc::ops
};

struct osm_b200_plan {
  PlanDesc d;
  int device = 0;
  int numSMs = 0;
  std::vector<StreamRt> st;
  std::vector<OpRt> ops;         // standalone ops (not fused into lld_kernel)
  PostParams pp;
  SeqPostParams sp;              // groups behind a Viterbi-smoothed pitch level (seq_post_kernel)
  int seqLagOp = -1;             // index into `ops` of the pitch chain whose lag they follow
  int *dErr = nullptr;           // device flag: a kernel left its supported geometry (checked after run_host)
  // the pitch chain (shs -> viterbi -> jitter: latency-bound, low occupancy) runs on its own stream next to the
  // other standalone ops of a step
  cudaStream_t auxStream = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  bool staticDirect = false;     // static rows are written straight into the output rows
  int identityOutCol = 0;
  bool fused = false;            // delta / delta-delta evaluated inside lld_kernel
  // batch bookkeeping
  std::vector<int64_t> cachedUttOff;
  PinBuf<long long> hMeta;       // uttOff | rowOff | statOff
  DevBuf<long long> dMeta;
  PinBuf<TileRef> hPost; DevBuf<TileRef> dPost; size_t nPostTiles = 0;
  std::vector<int32_t> uttPost0;
  DevBuf<float> dStat;
  DevBuf<float> dMeans;          // [nUtt][nStatic] column means for cFullinputMean groups
  bool needMeans = false;
  long long totalRows = 0, totalStat = 0, totalSamples = 0;
  size_t totalWork = 0;
  cudaEvent_t evMetaDone = nullptr, evK0 = nullptr, evKm = nullptr, evK1 = nullptr;
  bool metaPending = false, timed = false;
  // run_host buffers
  DevBuf<int16_t> dPcm;
  DevBuf<float> dPcmF;             // mono float samples of the batch (inputs in another format than 16-bit integer)
  DevBuf<float> dOut;
  cudaStream_t hostStream = nullptr, h2dStream = nullptr, d2hStream = nullptr;
  std::vector<cudaEvent_t> evPiece;   // 2 per pipeline piece: PCM landed / rows computed
  int lastLaunches = 0;
  // per-kernel profiling mode (osm_b200_plan_set_profiling): one event after every launch, everything on one stream
  bool profile = false;
  std::vector<cudaEvent_t> profEv;
  std::vector<const char *> profName;
  int profN = 0;
  LldLaunchInfo lastInfo{};
};
extern "C" {

int32_t osm_b200_abi_version(void) { return OSM_B200_ABI_VERSION; }
int32_t osm_b200_sizeof_component(void) { return (int32_t)sizeof(osm_b200_component); }

const char *osm_b200_last_error(void) { return g_err.c_str(); }

int32_t osm_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// defaults = the reference's ConfigType defaults
osm_b200_status osm_b200_component_defaults(int32_t type, osm_b200_component *c)
{
  if (!c || type < 0 || type >= OSM_B200_C_COUNT_) return fail(OSM_B200_ERR_INVALID, "bad component type");
  memset(c, 0, sizeof *c);
  c->type = type;
  c->copyInputName = 1;
  switch (type) {
    case OSM_B200_C_WAVESOURCE:
      c->u.wavesource.sampleRate = 16000; c->u.wavesource.nChannels = 1; c->u.wavesource.monoMixdown = 1;
      c->u.wavesource.format = OSM_B200_PCM_S16; strcpy(c->u.wavesource.outFieldName, "pcm");
      break;
    case OSM_B200_C_FRAMER:
      c->u.framer.frameSize = 0.025; c->u.framer.frameStep = 0.0;
      c->u.framer.frameCenterSpecialLeft = 1; c->u.framer.noPostEOIprocessing = 1;
      break;
    case OSM_B200_C_VECTORPREEMPHASIS: c->u.vectorpreemphasis.k = 0.97; c->u.vectorpreemphasis.de = 0; break;
    case OSM_B200_C_WINDOWER:
      c->u.windower.winFunc = OSM_B200_WIN_HANNING; c->u.windower.gain = 1.0; c->u.windower.offset = 0.0;
      c->u.windower.sigma = 0.4;
      c->u.windower.alpha0 = (1.0 - 0.16) * 0.5; c->u.windower.alpha1 = 0.5; c->u.windower.alpha2 = 0.16 * 0.5; c->u.windower.alpha3 = 0.0;
      c->u.windower.fade = 0.0; c->u.windower.squareRoot = 0;
      break;
    case OSM_B200_C_TRANSFORMFFT: c->u.transformfft.inverse = 0; c->u.transformfft.zeroPadSymmetric = 1; break;
    case OSM_B200_C_FFTMAGPHASE: c->u.fftmagphase.magnitude = 1; c->u.fftmagphase.dBpnorm = 90.302; c->u.fftmagphase.mindBp = -102.0; break;   // dspcore/fftmagphase.cpp:44-49
    case OSM_B200_C_MELSPEC:
      c->u.melspec.nBands = 26; c->u.melspec.lofreq = 20; c->u.melspec.hifreq = 8000;
      c->u.melspec.usePower = 0; c->u.melspec.htkcompatible = 1;
      break;
    case OSM_B200_C_MFCC:
      c->u.mfcc.firstMfcc = 1; c->u.mfcc.lastMfcc = 12; c->u.mfcc.melfloor = 1e-8; c->u.mfcc.doLog = 1;
      c->u.mfcc.cepLifter = 22; c->u.mfcc.htkcompatible = 1;
      break;
    case OSM_B200_C_PLP: {
      auto &p = c->u.plp;
      p.lpOrder = 5; p.nCeps = -1; p.firstCC = 1; p.lastCC = -1; p.doLog = 1; p.doAud = 1; p.RASTA = 0;
      p.newRASTA = 0; p.doInvLog = 1; p.doIDFT = 1; p.doLP = 1; p.doLpToCeps = 1; p.rastaUpperCutoff = 29;
      p.rastaLowerCutoff = 1; p.cepLifter = 0; p.compression = 0.33; p.melfloor = 9.3e-10; p.htkcompatible = 1;
      break;
    }
    case OSM_B200_C_SPECTRAL: {
      auto &s = c->u.spectral;
      s.squareInput = 1; s.flux = 1; s.centroid = 1; s.maxPos = 1; s.minPos = 1; s.oldSlopeScale = 1;
      s.specFloor = 1e-7;
      break;
    }
    case OSM_B200_C_ENERGY: {
      auto &e = c->u.energy;
      e.rms = 1; e.log = 1; e.escaleLog = e.escaleRms = e.escaleSquare = 1.0;
      break;
    }
    case OSM_B200_C_MZCR: c->u.mzcr.zcr = 1; c->u.mzcr.mcr = 1; c->u.mzcr.amax = 1; c->u.mzcr.maxmin = 1; break;
    case OSM_B200_C_ACF: {
      auto &a = c->u.acf;
      a.usePower = 1; a.expBeforeAbs = 1; a.symmetricData = 1; a.acfCepsNormOutput = 1;
      break;
    }
    case OSM_B200_C_PITCHACF: c->u.pitchacf.maxPitch = 500; c->u.pitchacf.voiceProb = 1; c->u.pitchacf.voicingCutoff = 0.55; break;
    case OSM_B200_C_DELTAREGRESSION: c->u.deltaregression.deltawin = 2; c->u.deltaregression.zeroSegBound = 1; break;
    case OSM_B200_C_CONTOURSMOOTHER: c->u.contoursmoother.smaWin = 3; break;
    case OSM_B200_C_INTENSITY: c->u.intensity.intensity = 1; c->u.intensity.loudness = 0; break;
    case OSM_B200_C_VECTORCONCAT: c->u.vectorconcat.processArrayFields = 1; c->u.vectorconcat.includeSingleElementFields = 0; break;
    case OSM_B200_C_SPECSCALE: {           // dsp/specScale.cpp:38-62 (scale "log" with logScaleBase 2 == octave)
      auto &q = c->u.specscale;
      q.scaleOctave = 1; q.sourceLin = 1; q.splineInterp = 1; q.minF = 25.0; q.maxF = -1.0;
      break;
    }
    case OSM_B200_C_PITCHSHS: {            // lldcore/pitchBase.cpp:41-62, lld/pitchShs.cpp:56-64
      auto &q = c->u.pitchshs;
      q.maxPitch = 620.0; q.minPitch = 52.0; q.nCandidates = 3; q.scores = 1; q.voicing = 1; q.voicingCutoff = 0.70;
      q.nHarmonics = 15; q.compressionFactor = 0.85;
      break;
    }
    case OSM_B200_C_PITCHSMOOTHERVITERBI: { // lld/pitchSmootherViterbi.cpp:45-68
      auto &q = c->u.pitchsmootherviterbi;
      q.bufferLength = 30; q.F0final = 1; q.wLocal = 2.0; q.wTvv = 10.0; q.wTvvd = 5.0; q.wTvuv = 10.0; q.wThr = 4.0;
      q.wRange = 1.0; q.wTuu = 0.0;
      break;
    }
    case OSM_B200_C_VALBASEDSELECTOR: c->u.valbasedselector.threshold = 1.0; break;   // other/valbasedSelector.cpp:35-49
    case OSM_B200_C_PITCHJITTER: {         // lld/pitchJitter.cpp:45-78
      auto &q = c->u.pitchjitter;
      snprintf(q.F0field, sizeof q.F0field, "%s", "F0final");
      q.searchRangeRel = 0.10; q.lgHNRfloor = -100.0; q.minNumPeriods = 2; q.minCC = 0.5; q.useBrokenJitterThresh = 1;
      break;
    }
    case OSM_B200_C_SPECRESAMPLE: c->u.specresample.targetFs = 16000.0; c->u.specresample.resampleRatio = -1.0; break;   // dsp/specResample.cpp:40-41
    case OSM_B200_C_LPC: c->u.lpc.p = 8; c->u.lpc.saveLPCoeff = 1; break;                                               // lld/lpc.cpp:33-45
    case OSM_B200_C_DATASELECTOR: c->u.dataselector.elementMode = 1; break;                                              // core/dataSelector.cpp:39
    case OSM_B200_C_HARMONICS: {           // lld/harmonics.cpp:28-56
      auto &q = c->u.harmonics;
      snprintf(q.f0ElementName, sizeof q.f0ElementName, "%s", "F0final");
      snprintf(q.magSpecFieldName, sizeof q.magSpecFieldName, "%s", "pcm_fftMag");
      q.f0ElementNameIsFull = 1; q.formantFrequencyFieldNameIsFull = 1; q.formantBandwidthFieldNameIsFull = 1;
      q.nHarmonics = 100; q.firstHarmonicMagnitude = 1; q.outputLogRelMagnitudes = 1; q.harmonicDifferencesLog = 1;
      q.formantAmplitudesLogRel = 1; q.formantAmplitudesStart = 1; q.formantAmplitudesEnd = -1; q.logRelValueFloorUnvoiced = -201.0;
      break;
    }
    case OSM_B200_C_FORMANTLPC: {          // lld/formantLpc.cpp:40-52
      auto &q = c->u.formantlpc;
      q.nFormants = -1; q.saveFormants = 1; q.minF = 50.0; q.maxF = 5500.0;
      break;
    }
    default: break;
  }
  return OSM_B200_OK;
}

static void build_twiddles(int M, std::vector<float2> &tw, int twOff[4])
{
  // factorisation must match Fact<M> in kernels.cu
  int R[3] = {0, 0, 0}, ns = 0;
  if (M == 256) { R[0] = 16; R[1] = 16; ns = 2; }
  else if (M == 512) { R[0] = 8; R[1] = 8; R[2] = 8; ns = 3; }
  else if (M == 1024) { R[0] = 16; R[1] = 16; R[2] = 4; ns = 3; }
  else { R[0] = 16; R[1] = 16; R[2] = 8; ns = 3; }
  int MS = M;
  tw.clear();
  for (int s = 0; s < 4; s++) twOff[s] = 0;
  for (int s = 0; s < ns - 1; s++) {       // the last stage has no twiddles
    const int stride = MS / R[s];
    twOff[s] = (int)tw.size();
    for (int j = 0; j < stride; j++)
      for (int q = 0; q < R[s]; q++) {
        const double ang = -2.0 * M_PI * (double)j * (double)q / (double)MS;
        tw.push_back(make_float2((float)cos(ang), (float)sin(ang)));
      }
    MS = stride;
  }
}


// pack the constant tables of one lld_kernel pass of a stream (front end + optional band op `opIdx`)
// and fill its LldParams.  `dumps`: this pass writes the magnitude level.
static osm_b200_status build_pass(osm_b200_plan *pl, int si, int opIdx, bool dumps, const cudaDeviceProp &prop,
                                  LldParams &kp, unsigned char *&dConstOut, PassRt &pr)
{
  const PlanDesc &d = pl->d;
  const Stream &sd = d.streams[si];
  StreamRt &rt = pl->st[si];
  const FrontEnd &fe = sd.fe;
  rt.runLld = sd.hasFft && (sd.fusedOp >= 0 || sd.dumpMag);
  rt.tileF = lld_supported_fft(fe.nfft) ? lld_tile_frames(fe.nfft) : 32;
  pr.op = opIdx;
  memset(&kp, 0, sizeof kp);
  kp.opKind = -1;
  kp.nChan = kernel_nchan(fe); kp.pcmF32 = fe.format != OSM_B200_PCM_S16;
  kp.frameSize = fe.frameSize; kp.frameStep = fe.frameStep;
  kp.hopMagic = (unsigned)((0x100000000ull + (unsigned long long)fe.frameStep - 1) / (unsigned long long)fe.frameStep);
  // per-lane stride S = frameStep + sPad of the shared-memory sample tile: odd S (scalar loads)
  // or even S with S/2 odd (64-bit sample-pair loads) is bank-conflict free across the lanes
  kp.sPad = (fe.frameStep % 2 != 0) ? 0 : (((fe.frameStep / 2) % 2 != 0) ? 0 : 2);
  kp.preemph = fe.preemph; kp.preDe = fe.preDe; kp.preK = fe.preK;
  kp.oneMinusK = 1 - fe.preK;                     // (1-k), float arithmetic (vectorPreemphasis.cpp:94)
  kp.winOffset = fe.winOffset; kp.hasWinOffset = fe.winOffset != 0.f;

  auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
  size_t o = 0;
  std::vector<unsigned char> blob;
  auto put = [&](const void *src, size_t bytes) { const size_t at = o; blob.resize(up16(o + bytes), 0); if (bytes) memcpy(&blob[at], src, bytes); o = up16(o + bytes); return at; };
  const size_t oWindow = put(fe.window.data(), fe.window.size() * sizeof(float));
  size_t oWin = 0, oTw = 0, oSplit = 0, oCoef = 0, oRange = 0, oDct = 0, oLift = 0, oEql = 0, oVisit = 0, oVB = 0;
  if (rt.runLld) {
    if (!lld_supported_fft(fe.nfft))
      return fail(OSM_B200_ERR_UNSUPPORTED, "FFT size " + std::to_string(fe.nfft) + " not supported (512, 1024, 2048, 4096)");
    const int M = fe.nfft / 2;
    std::vector<float4> winLut(M, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int e = 0; e < M; e++) {
      const int n = 2 * e;
      if (n < fe.frameSize) winLut[e].x = fe.window[n];
      if (n + 1 < fe.frameSize) winLut[e].y = fe.window[n + 1];
      const int off = (n < fe.frameSize) ? n + (n / fe.frameStep) * kp.sPad : 0;
      memcpy(&winLut[e].z, &off, sizeof(int));
      winLut[e].w = (n + 1 < fe.frameSize) ? 2.f : ((n < fe.frameSize) ? 1.f : 0.f);
    }
    std::vector<float2> tw;
    build_twiddles(M, tw, kp.twOff);
    kp.twCount = (int)tw.size();
    std::vector<float2> split(M / 2 + 1);
    for (int k = 0; k <= M / 2; k++) {
      const double ang = -2.0 * M_PI * (double)k / (double)fe.nfft;
      split[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    oWin = put(winLut.data(), winLut.size() * sizeof(float4));
    tw.push_back(make_float2(0.f, 0.f));
    oTw = put(tw.data(), tw.size() * sizeof(float2));
    oSplit = put(split.data(), split.size() * sizeof(float2));
    if (opIdx >= 0) {
      const StaticOp &op = d.ops[opIdx];
      const bool isPlp = op.kind == SOP_PLP;
      const MelBank &mb = d.mels[isPlp ? op.plp.melIdx : op.mfcc.melIdx];
      const MfccOp &mf = op.mfcc;
      const PlpOp &po = op.plp;
      if (isPlp && po.doLpToCeps && po.firstCC > 1) return fail(OSM_B200_ERR_UNSUPPORTED, "cPlp: firstCC > 1 is not supported");
      kp.opKind = isPlp ? 1 : 0;
      kp.nBands = mb.nBands; kp.melUsePower = mb.usePower;
      // without a magnitude dump the kernel keeps 2X (4|X|^2) out of the real-FFT split and the
      // exact factor 1/4 of the power path is folded into the band scale
      kp.melScale = (mb.usePower && !dumps) ? mb.outScale * 0.25f : mb.outScale;
      if (!isPlp) {
        kp.nStat = mf.nMfcc; kp.melfloor = mf.melfloor; kp.logMelfloor = mf.logMelfloor; kp.doLog = mf.doLog;
        kp.dctStride = (mb.nBands + 3) / 4 * 4; kp.dctRows = mf.nMfcc;
      } else {
        kp.nStat = po.nOut; kp.melfloor = po.melfloor; kp.logMelfloor = po.logMelfloor; kp.doLog = po.doLog;
        kp.plpAud = po.doAud; kp.plpInvLog = po.doInvLog; kp.plpIDFT = po.doIDFT; kp.plpLP = po.doLP; kp.plpCeps = po.doLpToCeps;
        kp.plpHtk = po.htk; kp.plpLifter = po.lifter; kp.plpOrder = po.lpOrder; kp.plpNAuto = po.nAuto; kp.plpNFreq = po.nFreq;
        kp.plpFirstCC = po.firstCC; kp.plpLastCC = po.lastCC; kp.plpCompression = po.compression;
        kp.dctStride = po.nFreq; kp.dctRows = po.nAuto;
      }
      // split the bands over the virtual warps: contiguous band groups, minimising the most expensive
      // group (cost model from the kernel's SASS: ~6 instructions per visited bin, ~50 per band for
      // scale / floor / log / store, ~12 per group) -- all warps meet at a barrier after this phase
      {
        const int nvw = lld_virtual_warps(fe.nfft);
        const int nB = mb.nBands;
        // the 512-point MFCC instance (lld_fast.cu) also adds every finished band into its DCT partial sums (~20 more)
        const double bandCost = (!isPlp && !dumps && lld_fast_applies(kp, fe.nfft)) ? 70.0 : 50.0;
        auto cost = [&](int bs, int be) -> double {          // bands [bs, be) visit ranges bs..be
          if (be <= bs) return 0.0;
          return 6.0 * (mb.rangeBegin[be + 1] - mb.rangeBegin[bs]) + bandCost * (be - bs) + 12.0;
        };
        // best[w][b] = minimal max-cost of covering bands [0, b) with w groups
        std::vector<std::vector<double>> best(nvw + 1, std::vector<double>(nB + 1, 1e30));
        std::vector<std::vector<int>> from(nvw + 1, std::vector<int>(nB + 1, 0));
        best[0][0] = 0.0;
        for (int w = 1; w <= nvw; w++)
          for (int b = 0; b <= nB; b++)
            for (int a = 0; a <= b; a++) {
              const double c = std::max(best[w - 1][a], cost(a, b));
              if (c < best[w][b]) { best[w][b] = c; from[w][b] = a; }
            }
        int b = nB;
        kp.melSplit[nvw] = nB;
        for (int w = nvw; w >= 1; w--) { b = from[w][b]; kp.melSplit[w - 1] = b; }
        for (int w = nvw + 1; w <= kMaxVW; w++) kp.melSplit[w] = nB;
      }
      std::vector<float> dctPad((size_t)kp.dctRows * kp.dctStride, 0.f);
      if (!isPlp) {
        for (int i = 0; i < mf.nMfcc; i++)
          memcpy(&dctPad[(size_t)i * kp.dctStride], &mf.cosT[(size_t)i * mb.nBands], sizeof(float) * mb.nBands);
      } else {
        memcpy(dctPad.data(), po.cosT.data(), sizeof(float) * po.cosT.size());
      }
      const std::vector<float> &liftV = isPlp ? po.lift : mf.liftFactor;
      std::vector<float> eqlV = isPlp ? po.eql : std::vector<float>(1, 0.f);
      oCoef = put(mb.coef.data(), mb.coef.size() * sizeof(float));
      oRange = put(mb.rangeBegin.data(), mb.rangeBegin.size() * sizeof(int));
      {
        std::vector<float2> visit;
        std::vector<int> vb(mb.nBands + 2, 0);
        for (int r = 0; r <= mb.nBands; r++) {
          vb[r] = (int)visit.size();
          for (int n = mb.rangeBegin[r]; n < mb.rangeBegin[r + 1]; n++) visit.push_back(make_float2(mb.coef[n], 1.0f - mb.coef[n]));
          while (visit.size() % 4) visit.push_back(make_float2(0.f, 0.f));
        }
        vb[mb.nBands + 1] = (int)visit.size();
        kp.melVCount = (int)visit.size();
        visit.push_back(make_float2(0.f, 0.f));
        oVisit = put(visit.data(), visit.size() * sizeof(float2));
        oVB = put(vb.data(), vb.size() * sizeof(int));
      }
      oDct = put(dctPad.data(), dctPad.size() * sizeof(float));
      oLift = put(liftV.data(), liftV.size() * sizeof(float));
      oEql = put(eqlV.data(), eqlV.size() * sizeof(float));
    }
  }
  unsigned char *dC = nullptr;
  CU(cudaMalloc(&dC, blob.size()));
  dConstOut = dC;
  CU(cudaMemcpy(dC, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  if (!rt.dWindow) rt.dWindow = reinterpret_cast<const float *>(dC + oWindow);
  if (rt.runLld) {
    kp.winLut = reinterpret_cast<const float4 *>(dC + oWin);
    kp.twiddles = reinterpret_cast<const float2 *>(dC + oTw);
    kp.splitTw = reinterpret_cast<const float2 *>(dC + oSplit);
    if (opIdx >= 0) {
      kp.melCoef = reinterpret_cast<const float *>(dC + oCoef);
      kp.melRange = reinterpret_cast<const int *>(dC + oRange);
      kp.melVisit = reinterpret_cast<const float2 *>(dC + oVisit);
      kp.melVB = reinterpret_cast<const int *>(dC + oVB);
      kp.dctCos = reinterpret_cast<const float *>(dC + oDct);
      kp.dctLift = reinterpret_cast<const float *>(dC + oLift);
      kp.plpEql = reinterpret_cast<const float *>(dC + oEql);
      const StaticOp &op = d.ops[opIdx];
      if (op.kind == SOP_PLP && op.plp.rasta) {
        // RASTA sits in the middle of cPlp: the kernel pass stops at the (log) band level, a
        // sequential filter runs over time, plp_tail_kernel finishes the op
        pr.rasta = true;
        pr.tail = kp;
        kp.plpAud = 0; kp.plpInvLog = 0; kp.plpIDFT = 0; kp.plpLP = 0; kp.plpCeps = 0; kp.plpLifter = 0;
        kp.nStat = kp.nBands;
        memset(&pr.rp, 0, sizeof pr.rp);
        pr.rp.nBands = kp.nBands; pr.rp.frameSize = fe.frameSize; pr.rp.frameStep = fe.frameStep;
        pr.rp.mode = op.plp.rasta; pr.rp.iir = op.plp.rastaIir;
        for (int i = 0; i < 5; i++) pr.rp.fir[i] = op.plp.rastaFir[i];
      }
    }
    if (rt.forceNarrow || (lld_smem_bytes(kp, fe.nfft) > (size_t)prop.sharedMemPerBlockOptin && fe.nfft >= 1024)) {
      kp.narrow = 1;                                   // long stereo strides: half-width tiles
      rt.tileF = lld_tile_frames(fe.nfft, true);
      rt.forceNarrow = true;                           // every pass of a stream uses the same tile width
    }
    if (lld_smem_bytes(kp, fe.nfft) > (size_t)prop.sharedMemPerBlockOptin)
      return fail(OSM_B200_ERR_UNSUPPORTED, "configuration needs more shared memory than the device offers");
  }
  return OSM_B200_OK;
}

osm_b200_status osm_b200_plan_create(const osm_b200_component *comps, int32_t n_comps,
                                     const char *output_level, int32_t device, osm_b200_plan **out)
{
  if (!comps || n_comps <= 0 || !out) return fail(OSM_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  osm_b200_plan *pl = new (std::nothrow) osm_b200_plan();
  if (!pl) return fail(OSM_B200_ERR_NOMEM, "out of memory");
  std::string err;
  osm_b200_status st = compile_graph(comps, n_comps, output_level, pl->d, err);
  if (st != OSM_B200_OK) { delete pl; return fail(st, err); }
  const PlanDesc &d = pl->d;
  for (const Stream &sd : d.streams)
    if (sd.hasFft && (sd.fusedOp >= 0 || sd.dumpMag) && !lld_supported_fft(sd.fe.nfft)) {
      const int nf = sd.fe.nfft;
      delete pl;
      return fail(OSM_B200_ERR_UNSUPPORTED, "FFT size " + std::to_string(nf) + " not supported (512, 1024, 2048, 4096)");
    }
  if (device < 0) {
    // description-only plan: geometry, names and frame-count rules without touching CUDA
    // (used by host-side tooling and CPU-only tests); it can never run.
    pl->device = -1;
    *out = pl;
    return OSM_B200_OK;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    delete pl;
    return fail(OSM_B200_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
  }
  if (device >= ndev) { delete pl; return fail(OSM_B200_ERR_INVALID, "bad device index"); }
  pl->device = device;
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { osm_b200_plan_destroy(pl); return cuda_fail(e_, #call); } } while (0)
#define STP(call) do { osm_b200_status s_ = (call); if (s_ != OSM_B200_OK) { osm_b200_plan_destroy(pl); return s_; } } while (0)
  CUP(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUP(cudaGetDeviceProperties(&prop, device));
  pl->numSMs = prop.multiProcessorCount;

  pl->st.resize(d.streams.size());
  for (size_t s = 0; s < d.streams.size(); s++) {
    StreamRt &rt = pl->st[s];
    const std::vector<int> &bo = d.streams[s].bandOps;
    // pass 0: first band op (or none) + the magnitude dump; passes 1..: the other band ops
    STP(build_pass(pl, (int)s, bo.empty() ? -1 : bo[0], d.streams[s].dumpMag, prop, rt.kp, rt.dConst, rt.pass0x));
    rt.extra.resize(bo.size() > 1 ? bo.size() - 1 : 0);
    for (size_t j = 1; j < bo.size(); j++) {
      PassRt &pr = rt.extra[j - 1];
      STP(build_pass(pl, (int)s, bo[j], false, prop, pr.kp, pr.dConst, pr));
    }
    if (rt.forceNarrow) {                              // a later pass switched to half-width tiles: all must agree
      rt.kp.narrow = 1;
      for (PassRt &pr : rt.extra) pr.kp.narrow = 1;
    }
  }

  // ---- output groups / execution mode ----
  PostParams &pp = pl->pp;
  memset(&pp, 0, sizeof pp);
  // "simple" plans: one stream, one fused band op, nothing else.  Their static rows go straight
  // into the output rows; delta / delta-delta can then be fused into lld_kernel as well.
  const bool simple = d.streams.size() == 1 && d.ops.size() == 1 && d.streams[0].fusedOp == 0 && !d.streams[0].dumpMag &&
                      !(d.ops[0].kind == SOP_PLP && d.ops[0].plp.rasta);
  pl->staticDirect = false;
  if (simple)
    for (const auto &g : d.groups)
      if (g.stages.empty() && g.srcCol == 0 && g.n == d.nStatic) { pl->staticDirect = true; pl->identityOutCol = g.outCol; break; }
  memset(&pl->sp, 0, sizeof pl->sp);
  int seqLagDescOp = -1;
  for (const auto &g : d.groups) {
    if (g.stages.empty() && pl->staticDirect && g.srcCol == 0 && g.n == d.nStatic && g.outCol == pl->identityOutCol) continue;
    if (g.lagKind != 0 && !g.stages.empty()) {
      SeqPostParams &sq = pl->sp;
      if (sq.nGroups >= kMaxSeqGroups) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "too many output groups behind the pitch chain"); }
      if (seqLagDescOp >= 0 && seqLagDescOp != g.lagOp) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "more than one SHS pitch chain per graph is not supported"); }
      seqLagDescOp = g.lagOp;
      SeqGroup &G = sq.groups[sq.nGroups++];
      G.srcCol = g.srcCol; G.n = g.n; G.outCol = g.outCol; G.lagKind = g.lagKind; G.nStages = (int)g.stages.size();
      G.noZero = g.stages[0].flags & 1;
      G.deltaWin = g.stages.size() > 1 ? g.stages[1].win : 0;
      G.segId = g.segId;
      G.gateCol = g.gateCol; G.gateFlags = (g.gateInvert ? 1 : 0) | (g.gateAllowEqual ? 2 : 0); G.gateThr = g.gateThreshold; G.gateOut = g.gateOutVal;
      if (G.nStages == 2 && G.deltaWin != 2) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cDeltaRegression(onlyInSegments) behind the pitch chain: deltawin must be 2"); }
      if (G.nStages == 2) {
        int segCols = 0;
        for (int q = 0; q < sq.nGroups; q++) if (sq.groups[q].nStages == 2 && sq.groups[q].segId == G.segId) segCols += sq.groups[q].n;
        if (segCols > 32) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cDeltaRegression(onlyInSegments): more than 32 elements"); }
      }
      if (G.segId >= kMaxSegIds) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "too many onlyInSegments delta components"); }
      sq.frameSize = d.streams[g.stream].fe.frameSize; sq.frameStep = d.streams[g.stream].fe.frameStep;
      continue;
    }
    if (pp.nGroups >= kMaxPostGroups) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "too many output groups"); }
    PostGroup &pg = pp.groups[pp.nGroups++];
    pg.srcCol = g.srcCol; pg.n = g.n; pg.outCol = g.outCol; pg.nStages = (int)g.stages.size();
    pg.frameSize = d.streams[g.stream].fe.frameSize; pg.frameStep = d.streams[g.stream].fe.frameStep;
    pg.nLim = 0;
    for (int ls : g.limitStreams) if (pg.nLim < 3) { pg.limSize[pg.nLim] = d.streams[ls].fe.frameSize; pg.limStep[pg.nLim] = d.streams[ls].fe.frameStep; pg.nLim++; }
    for (size_t i = 0; i < g.stages.size(); i++) { pg.kind[i] = g.stages[i].kind; pg.win[i] = g.stages[i].win; pg.flags[i] = g.stages[i].flags; if (g.stages[i].kind == ST_CMS) pl->needMeans = true; }
  }
  // fused pattern: [static | delta(W1) | delta(W1,W2)] over the whole static vector
  pl->fused = false;
  {
    LldParams &kp = pl->st[0].kp;
    kp.fused = 0; kp.halo = 0; kp.fW1 = kp.fW2 = 0; kp.fNorm1 = kp.fNorm2 = 1.f;
    if (simple && pl->staticDirect && pl->identityOutCol == 0 && d.groups.size() == 3 && d.nOut == 3 * d.nStatic) {
      const auto &g1 = d.groups[1], &g2 = d.groups[2];
      const bool shape = d.groups[0].stages.empty() && g1.stages.size() == 1 && g2.stages.size() == 2 &&
                         g1.srcCol == 0 && g2.srcCol == 0 && g1.n == d.nStatic && g2.n == d.nStatic &&
                         g1.outCol == d.nStatic && g2.outCol == 2 * d.nStatic &&
                         g1.stages[0].kind == ST_DELTA && g2.stages[0].kind == ST_DELTA && g2.stages[1].kind == ST_DELTA &&
                         g1.stages[0].win == g2.stages[0].win &&
                         g1.stages[0].flags == 0 && g2.stages[0].flags == 0 && g2.stages[1].flags == 0;   // plain regression only
      // OSM_B200_NO_FUSE=1 forces the two-kernel path (used by the tests to cross-check both)
      const char *nf = getenv("OSM_B200_NO_FUSE");
      // the kernel keeps the statics of two tiles (2F frames): a row needs 2*halo+1 of them
      const int hl = g1.stages[0].win + g2.stages[1].win, tF = pl->st[0].tileF;
      if (shape && hl <= 8 && 2 * hl + 1 <= 2 * tF && !(nf && nf[0] == '1')) {
        pl->fused = true;
        kp.fused = 1; kp.fW1 = g1.stages[0].win; kp.fW2 = g2.stages[1].win; kp.halo = kp.fW1 + kp.fW2;
        auto normOf = [](int W) { float n = 0.f; for (int i = 1; i <= W; i++) n += (float)i * (float)i; return n * 2.0f; };
        kp.fNorm1 = normOf(kp.fW1); kp.fNorm2 = normOf(kp.fW2);    // deltaRegression.cpp:77-80
        // divisors with an exhaustively verified exact reciprocal+FMA division (kernels.cu div_exact)
        auto rcpOf = [](float n) { return (n == 2.f || n == 10.f || n == 28.f || n == 60.f) ? 1.0f / n : 0.f; };
        kp.fRcp1 = rcpOf(kp.fNorm1); kp.fRcp2 = rcpOf(kp.fNorm2);
      }
    }
  }
  pp.nStat = d.nStatic; pp.maxN = 1; pp.halo = 0;
  for (int g = 0; g < pp.nGroups; g++) {
    int h = 0;
    for (int s2 = 0; s2 < pp.groups[g].nStages; s2++) h += pp.groups[g].win[s2];
    if (h > pp.halo) pp.halo = h;
    if (pp.groups[g].n > pp.maxN) pp.maxN = pp.groups[g].n;
  }
  if (pp.halo > 12) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "summed temporal half windows exceed 12 frames"); }
  pp.rows = post_tile_rows(pp.nStat, pp.maxN, pp.halo);
  if ((size_t)(pp.rows + 2 * pp.halo) * (pp.nStat + 2 * pp.maxN) * sizeof(float) > 200 * 1024) {
    osm_b200_plan_destroy(pl);
    return fail(OSM_B200_ERR_UNSUPPORTED, "the static level is too wide for the row assembly kernel");
  }

  // ---- standalone ops ----
  for (size_t oi = 0; oi < d.ops.size(); oi++) {
    const StaticOp &op = d.ops[oi];
    if (op.kind == SOP_MFCC || op.kind == SOP_PLP) continue;      // fused into its stream's lld_kernel
    OpRt rt;
    rt.kind = op.kind; rt.stream = op.stream; rt.descOp = (int)oi;
    if (op.kind == SOP_MAG) {
      pl->st[op.stream].needTiles = true;
      rt.vN = d.streams[op.stream].fe.nBins; rt.vOutCol = op.outCol;
      rt.magMode = op.magMode; rt.magN = (float)d.streams[op.stream].fe.nfft; rt.magDbNorm = op.magDbNorm; rt.magMinDb = op.magMinDb;
      pl->ops.push_back(rt);
      continue;
    }
    if (op.kind == SOP_VECOP) {
      const StaticOp &src = d.ops[op.srcOp];
      if (src.kind != SOP_MFCC && src.kind != SOP_PLP) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cVectorOperation: the input must be a cMfcc / cPlp level"); }
      rt.vSrcCol = src.outCol; rt.vN = src.nOut; rt.vOutCol = op.outCol;
      pl->ops.push_back(rt);
      continue;
    }
    const FrontEnd &fe = d.streams[op.stream].fe;
    StreamRt &srt = pl->st[op.stream];
    srt.needTiles = true;
    if (op.kind == SOP_SPECTRAL) {
      const SpectralOp &so = op.spectral;
      SpectralParams &sp = rt.sp;
      memset(&sp, 0, sizeof sp);
      sp.F = srt.tileF; sp.statStride = d.nStatic; sp.outCol = op.outCol;
      sp.nSrc = so.nSrc; sp.loBin = so.loBin; sp.hiBin = so.hiBin; sp.F0 = so.F0;
      sp.squareInput = so.squareInput; sp.useLog = so.useLog; sp.normBand = so.normBand; sp.buggyRollOff = so.buggyRollOff;
      sp.oldSlopeScale = so.oldSlopeScale; sp.reqMag = so.reqMag; sp.reqPow = so.reqPow; sp.reqLog = so.reqLog;
      sp.specFloor = so.specFloor; sp.logSpecFloor = so.logSpecFloor;
      sp.nBands = (int)so.bandIL.size();
      for (int i = 0; i < sp.nBands; i++) { sp.bandIL[i] = so.bandIL[i]; sp.bandIR[i] = so.bandIR[i]; sp.bandWL[i] = so.bandWL[i]; sp.bandWR[i] = so.bandWR[i]; }
      sp.nSlopes = (int)so.slopeIL.size();
      for (int i = 0; i < sp.nSlopes; i++) { sp.slopeIL[i] = so.slopeIL[i]; sp.slopeIR[i] = so.slopeIR[i]; sp.slopeWL[i] = so.slopeWL[i]; sp.slopeWR[i] = so.slopeWR[i]; sp.slopeNind[i] = so.slopeNind[i]; }
      sp.nRollOff = (int)so.rollOff.size();
      for (int i = 0; i < sp.nRollOff; i++) sp.rollOff[i] = so.rollOff[i];
      sp.alphaRatio = so.alphaRatio; sp.hammarberg = so.hammarberg; sp.flux = so.flux; sp.centroid = so.centroid;
      sp.maxPos = so.maxPos; sp.minPos = so.minPos; sp.entropy = so.entropy; sp.stddev = so.stddev; sp.variance = so.variance;
      sp.skewness = so.skewness; sp.kurtosis = so.kurtosis; sp.slope = so.slope; sp.sharpness = so.sharpness;
      sp.harmonicity = so.harmonicity; sp.flatness = so.flatness; sp.logFlatness = so.logFlatness;
      CUP(cudaMalloc(&rt.dSharpW, so.sharpW.size() * sizeof(double)));
      CUP(cudaMemcpy(rt.dSharpW, so.sharpW.data(), so.sharpW.size() * sizeof(double), cudaMemcpyHostToDevice));
      sp.sharpW = rt.dSharpW;
    } else if (op.kind == SOP_HARMONICS) {
      const HarmonicsOp &ho = op.harmonics;
      HarmonicsParams &hp = rt.hrm;
      memset(&hp, 0, sizeof hp);
      const int N = (ho.nb - 1) * 2;
      std::vector<double> ct(N);
      for (int m = 0; m < N; m++) ct[m] = cos(2.0 * M_PI * (double)m / (double)N);
      CUP(cudaMalloc(&rt.dCosTab, ct.size() * sizeof(double)));
      CUP(cudaMemcpy(rt.dCosTab, ct.data(), ct.size() * sizeof(double), cudaMemcpyHostToDevice));
      hp.F = srt.tileF; hp.nb = ho.nb; hp.binHz = ho.binHz; hp.statStride = d.nStatic; hp.outCol = op.outCol;
      hp.f0Col = ho.f0Col; hp.fmtCol = ho.fmtCol; hp.nFmt = ho.nFmt; hp.cosTab = rt.dCosTab;
      hp.nHarm = ho.nHarm; hp.doHnr = ho.hnr; hp.nDiffs = (int)ho.diffs.size() / 4;
      for (size_t i = 0; i < ho.diffs.size() && i < 16; i++) hp.diffs[i] = ho.diffs[i];
      hp.doFa = ho.fa; hp.faStart = ho.faStart; hp.faEnd = ho.faEnd; hp.floorUnvoiced = ho.floorUnvoiced;
      if (harmonics_smem_bytes(hp) > 200 * 1024) { pl->ops.push_back(rt); osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cHarmonics: spectrum too long for the kernel"); }
    } else if (op.kind == SOP_PITCH) {
      const PitchChainOp &pc = op.chain;
      // one blob: fwdA | fwdP6 | r1 | r2 | bwdD (double[nMag]) | ia | ic | id | audW (double[nPts]) | ik (int[nPts]) | shift (int) | hscale (float)
      // the five recurrence tables are stored lane-interleaved for the blocked scan of shs_kernel (lane l owns points
      // [l * blk, (l+1) * blk)): entry of (lane, step k) at [k * 32 + lane], so a warp reads 32 consecutive doubles
      const int blk = (pc.nMag + 31) / 32;
      const size_t nM = (size_t)blk * 32, nP = (size_t)pc.nPts, nH = pc.shift.size();
      std::vector<unsigned char> blob((5 * nM + 4 * nP) * sizeof(double) + nP * sizeof(int) + nH * (sizeof(int) + sizeof(float)) + 64);
      double *bd = reinterpret_cast<double *>(blob.data());
      const std::vector<double> *tabs[5] = {&pc.fwdA, &pc.fwdP6, &pc.r1, &pc.r2, &pc.bwdD};
      for (int t = 0; t < 5; t++)
        for (int l = 0; l < 32; l++)
          for (int k = 0; k < blk; k++) {
            const int i = l * blk + k;
            bd[(size_t)t * nM + (size_t)k * 32 + l] = i < pc.nMag ? (*tabs[t])[i] : 0.0;
          }
      double *bp = bd + 5 * nM;
      memcpy(bp, pc.ia.data(), nP * 8); memcpy(bp + nP, pc.ic.data(), nP * 8); memcpy(bp + 2 * nP, pc.id.data(), nP * 8);
      if (!pc.audW.empty()) memcpy(bp + 3 * nP, pc.audW.data(), nP * 8);
      int *bi = reinterpret_cast<int *>(bp + 4 * nP);
      memcpy(bi, pc.ik.data(), nP * sizeof(int));
      memcpy(bi + nP, pc.shift.data(), nH * sizeof(int));
      memcpy(bi + nP + nH, pc.hscale.data(), nH * sizeof(float));
      CUP(cudaMalloc(&rt.dPitchTab, blob.size()));
      CUP(cudaMemcpy(rt.dPitchTab, blob.data(), blob.size(), cudaMemcpyHostToDevice));
      const double *dd = reinterpret_cast<const double *>(rt.dPitchTab);
      ShsParams &sh = rt.shs;
      memset(&sh, 0, sizeof sh);
      sh.F = srt.tileF; sh.nShsCols = pc.nShsCols; sh.nMag = pc.nMag; sh.nPts = pc.nPts; sh.blk = (pc.nMag + 31) / 32;
      sh.enhance = pc.enhance; sh.smooth = pc.smooth; sh.hasAudW = !pc.audW.empty();
      sh.fwdA = dd; sh.fwdP6 = dd + nM; sh.r1 = dd + 2 * nM; sh.r2 = dd + 3 * nM; sh.bwdD = dd + 4 * nM;
      const double *dp = dd + 5 * nM;
      sh.ia = dp; sh.ic = dp + nP; sh.id = dp + 2 * nP; sh.audW = dp + 3 * nP;
      const int *di = reinterpret_cast<const int *>(dp + 4 * nP);
      sh.ik = di;
      for (size_t h = 0; h < nH && h < 32; h++) { sh.shift[h] = pc.shift[h]; sh.hscale[h] = pc.hscale[h]; }
      sh.nCand = pc.nCand; sh.nHarm = pc.nHarm; sh.Fmint = pc.Fmint; sh.Fstept = pc.Fstept; sh.logBase = pc.logBase;
      sh.maxPitch = pc.maxPitch; sh.minPitch = pc.minPitch; sh.voicingCutoff = pc.voicingCutoff; sh.lfCutBin = pc.lfCutBin;
      sh.greedy = pc.greedy; sh.octaveCorr = pc.octaveCorr; sh.scores = pc.scores; sh.voicing = pc.voicing; sh.F0C1 = pc.F0C1;
      sh.voicingC1 = pc.voicingC1; sh.F0raw = pc.F0raw; sh.voicingClip = pc.voicingClip;
      if (((size_t)2 * (pc.nMag + 2) * sizeof(double) + (size_t)pc.nPts * sizeof(float) + 128) > (size_t)prop.sharedMemPerBlockOptin ||
          (size_t)pc.nPts * sizeof(float) > (size_t)(pc.nMag + 2) * sizeof(double)) {
        osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cSpecScale: spectrum too long for the SHS kernel's workspace");
      }
      ViterbiParams &vp = rt.vit;
      memset(&vp, 0, sizeof vp);
      vp.nShsCols = pc.nShsCols; vp.nCand = pc.nCand; vp.frameSize = fe.frameSize; vp.frameStep = fe.frameStep;
      vp.statStride = d.nStatic; vp.outCol = op.outCol; vp.bufLen = pc.bufLen;
      vp.oF0final = pc.oF0final; vp.oF0finalLog = pc.oF0finalLog; vp.oF0finalEnv = pc.oF0finalEnv; vp.oF0finalEnvLog = pc.oF0finalEnvLog;
      vp.oVClipped = pc.oVClipped; vp.oVUnclipped = pc.oVUnclipped;
      vp.wLocal = pc.wLocal; vp.wTvv = pc.wTvv; vp.wTvvd = pc.wTvvd; vp.wTvuv = pc.wTvuv; vp.wThr = pc.wThr; vp.wRange = pc.wRange; vp.wTuu = pc.wTuu;
      vp.voiceThresh = pc.voicingCutoff;                       // level meta data of cPitchShs (lldcore/pitchBase.cpp:150-153)
      vp.hasSel = pc.hasSel; vp.selCol = pc.hasSel ? d.ops[pc.selOp].outCol : 0; vp.selInvert = pc.selInvert; vp.selAllowEqual = pc.selAllowEqual;
      vp.selThreshold = pc.selThreshold; vp.selOutputVal = pc.selOutputVal;
      if (pc.nCand + 1 > 9) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cPitchShs.nCandidates > 8 is not supported"); }
    } else if (op.kind == SOP_JITTER) {
      const JitterOp &jo = op.jitter;
      JitterParams &jp = rt.jit;
      memset(&jp, 0, sizeof jp);
      jp.nChan = kernel_nchan(fe); jp.pcmF32 = fe.format != OSM_B200_PCM_S16; jp.frameSize = fe.frameSize; jp.frameStep = fe.frameStep;
      jp.Ts = 1.0 / fe.sampleRate;                               // period of the wave level
      jp.pitchT = fe.frameStepSec;                               // period of the F0 level (core/winToVecProcessor.cpp:563)
      jp.statStride = d.nStatic; jp.f0Col = d.ops[jo.pitchOp].outCol + jo.f0Col; jp.outCol = op.outCol;
      jp.searchRangeRel = jo.searchRangeRel; jp.lgHNRfloor = (float)jo.lgHNRfloor; jp.minNumPeriods = jo.minNumPeriods;
      float thr = (float)jo.minCC;                               // lld/pitchJitter.cpp:150-158
      if (thr < 0.01f) thr = 0.01f;
      if (thr > 0.99f) thr = 0.99f;
      jp.threshCC = thr;
      jp.jitterLocal = jo.jitterLocal; jp.jitterDDP = jo.jitterDDP; jp.jitterLocalEnv = jo.jitterLocalEnv; jp.jitterDDPEnv = jo.jitterDDPEnv;
      jp.shimmerLocal = jo.shimmerLocal; jp.shimmerLocalDB = jo.shimmerLocalDB; jp.shimmerLocalEnv = jo.shimmerLocalEnv;
      jp.shimmerLocalDBEnv = jo.shimmerLocalDBEnv; jp.harmonicERMS = jo.harmonicERMS; jp.noiseERMS = jo.noiseERMS; jp.linearHNR = jo.linearHNR;
      jp.logHNR = jo.logHNR; jp.shimmerUseRms = jo.shimmerUseRms; jp.refinedF0 = jo.refinedF0; jp.srcQualRange = jo.srcQualRange;
      jp.srcQualMean = jo.srcQualMean; jp.peakToPeak = jo.peakToPeak; jp.brokenThresh = jo.brokenThresh;
      {
        // F0 comes out of the pitch chain within [minPitch, maxPitch]: longest / shortest period in samples
        const PitchChainOp &pc = d.ops[jo.pitchOp].chain;
        const double fs = fe.sampleRate, fLo = pc.minPitch > 1.0 ? pc.minPitch : 1.0, fHi = pc.maxPitch > fLo ? pc.maxPitch : fLo;
        const double tMax = fs / fLo, tMin = fs / fHi;
        const long maxPer = (long)ceil((1.0 + jo.searchRangeRel) * tMax) + 2;
        long minPer = (long)floor((1.0 - jo.searchRangeRel) * tMin);
        if (minPer < 1) minPer = 1;
        const long ppLen = (long)ceil(fe.frameStepSec * fs) + 1;
        long capWav = fe.frameSize + 2 + ppLen + maxPer + 16;
        const long twoPp = jo.minNumPeriods * maxPer + jo.minNumPeriods + ppLen + 16;
        if (capWav < twoPp) capWav = twoPp;
        jp.capWav = (int)((capWav + 3) & ~3L);
        jp.capCC = (int)(((long)ceil(2.0 * jo.searchRangeRel * tMax) + 8 + 1) & ~1L);
        jp.capAvg = (int)(((long)ceil(tMax) + 8 + 3) & ~3L);
        jp.capPb = (int)((capWav / minPer + 8 + 3) & ~3L);
        if (pc.minPitch < 1.0 || (size_t)4 * ((size_t)jp.capCC * 8 + (size_t)(jp.capWav + jp.capAvg) * 4 + (size_t)jp.capPb * 4) > 200 * 1024) {
          osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cPitchJitter: frame size / pitch range need more workspace than the kernel has");
        }
      }
    } else if (op.kind == SOP_PITCHACF) {
      if (!acf_pitch_supported_fft(fe.nfft)) { osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cAcf / cPitchACF: FFT size must be 512, 1024 or 2048"); }
      const PitchAcfOp &po = op.pitch;
      AcfPitchParams &ap = rt.ap;
      memset(&ap, 0, sizeof ap);
      ap.F = srt.tileF;     // width of the magnitude level's tiles
      ap.nfft = fe.nfft; ap.nSrc = fe.nBins;
      ap.acfUsePower = po.acfUsePower; ap.cepUsePower = po.cepUsePower; ap.absCepstrum = po.absCepstrum; ap.normOutput = po.normOutput;
      ap.maxPitch = po.maxPitch; ap.voicingCutoff = po.voicingCutoff; ap.fsSec = po.fsSec;
      ap.voiceProb = po.voiceProb; ap.voiceQual = po.voiceQual; ap.HNR = po.HNR; ap.HNRdB = po.HNRdB; ap.linHNR = po.linHNR;
      ap.F0 = po.F0; ap.F0raw = po.F0raw; ap.F0env = po.F0env;
      ap.statStride = d.nStatic; ap.outCol = op.outCol; ap.frameSize = fe.frameSize; ap.frameStep = fe.frameStep;
      // complex FFT of size nfft: same factorisation tables as lld_kernel's M = nfft
      std::vector<float2> tw;
      build_twiddles(fe.nfft, tw, ap.twOff);
      ap.twCount = (int)tw.size();
      CUP(cudaMalloc(&rt.dTw, (tw.size() + 1) * sizeof(float2)));
      CUP(cudaMemcpy(rt.dTw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
      ap.twiddles = rt.dTw;
    } else {
      TimeOpParams &tp = rt.tp;
      memset(&tp, 0, sizeof tp);
      tp.nChan = kernel_nchan(fe); tp.pcmF32 = fe.format != OSM_B200_PCM_S16; tp.F = srt.tileF; tp.statStride = d.nStatic; tp.outCol = op.outCol;
      tp.frameSize = fe.frameSize; tp.frameStep = fe.frameStep;
      tp.windowed = op.windowed; tp.preemph = op.windowed && fe.preemph; tp.preDe = fe.preDe; tp.preK = fe.preK;
      tp.oneMinusK = 1 - fe.preK; tp.winOffset = fe.winOffset; tp.window = srt.dWindow;
      if (op.kind == SOP_INTENSITY) {
        tp.iIntensity = op.intensity.intensity; tp.iLoudness = op.intensity.loudness;
        tp.iW0 = op.intensity.w[0]; tp.iW1 = op.intensity.w[1]; tp.iWinSum = op.intensity.winSum;
      } else if (op.kind == SOP_FORMANT) {
        const FormantOp &fo = op.formant;
        FormantParams &fp = rt.fmt;
        memset(&fp, 0, sizeof fp);
        CUP(cudaMalloc(&rt.dFmtD, fo.D.size() * sizeof(float)));
        CUP(cudaMemcpy(rt.dFmtD, fo.D.data(), fo.D.size() * sizeof(float), cudaMemcpyHostToDevice));
        fp.tp = tp; fp.D = rt.dFmtD; fp.nRes = fo.nRes; fp.nResPad = fo.nResPad; fp.p = fo.p; fp.nFormants = fo.nFormants;
        fp.T = fo.T; fp.minF = fo.minF; fp.maxF = fo.maxF;
        fp.refOrder = fo.refOrder ? 1 : 0; fp.kHalf = fo.kHalf; fp.padLeft = fo.padLeft; fp.halfK = fo.halfK;
        fp.saveFormants = fo.saveFormants; fp.saveBandwidths = fo.saveBandwidths; fp.saveNValid = fo.saveNValid;
        if (formant_smem_bytes(fp) > 200 * 1024) { pl->ops.push_back(rt); osm_b200_plan_destroy(pl); return fail(OSM_B200_ERR_UNSUPPORTED, "cSpecResample: frame too long for the formant kernel"); }
      } else if (op.kind == SOP_ENERGY) {
        const EnergyOp &e = op.energy;
        tp.eHtk = e.htk; tp.eRms = e.rms; tp.eEnergy2 = e.energy2; tp.eLog = e.lg;
        tp.escaleLog = e.escaleLog; tp.escaleRms = e.escaleRms; tp.escaleSquare = e.escaleSquare;
        tp.ebiasLog = e.ebiasLog; tp.ebiasRms = e.ebiasRms; tp.ebiasSquare = e.ebiasSquare;
      } else {
        const MzcrOp &z = op.mzcr;
        tp.zZcr = z.zcr; tp.zMcr = z.mcr; tp.zAmax = z.amax; tp.zMaxmin = z.maxmin; tp.zDc = z.dc;
      }
    }
    pl->ops.push_back(rt);
  }

  if (seqLagDescOp >= 0) {
    CUP(cudaStreamCreateWithFlags(&pl->auxStream, cudaStreamNonBlocking));
    CUP(cudaEventCreateWithFlags(&pl->evFork, cudaEventDisableTiming));
    CUP(cudaEventCreateWithFlags(&pl->evJoin, cudaEventDisableTiming));
  }
  CUP(cudaMalloc(&pl->dErr, sizeof(int)));
  CUP(cudaMemset(pl->dErr, 0, sizeof(int)));
  if (seqLagDescOp >= 0)
    for (size_t i = 0; i < pl->ops.size(); i++) if (pl->ops[i].descOp == seqLagDescOp) pl->seqLagOp = (int)i;
  CUP(cudaEventCreateWithFlags(&pl->evMetaDone, cudaEventDisableTiming));
  CUP(cudaEventCreate(&pl->evK0));
  CUP(cudaEventCreate(&pl->evK1));
  CUP(cudaEventCreate(&pl->evKm));
#undef CUP
#undef STP
  *out = pl;
  return OSM_B200_OK;
}

void osm_b200_plan_destroy(osm_b200_plan *pl)
{
  if (!pl) return;
  if (pl->device < 0) { delete pl; return; }
  cudaSetDevice(pl->device);
  if (pl->hostStream) { cudaStreamSynchronize(pl->hostStream); cudaStreamDestroy(pl->hostStream); }
  if (pl->h2dStream) cudaStreamDestroy(pl->h2dStream);
  if (pl->d2hStream) cudaStreamDestroy(pl->d2hStream);
  for (cudaEvent_t e : pl->evPiece) cudaEventDestroy(e);
  for (cudaEvent_t e : pl->profEv) cudaEventDestroy(e);
  cudaDeviceSynchronize();
  for (StreamRt &s : pl->st) {
    if (s.dConst) cudaFree(s.dConst);
    s.pass0x.dBand.release();
    for (PassRt &pr : s.extra) { if (pr.dConst) cudaFree(pr.dConst); pr.dBand.release(); }
    s.hChunks.release(); s.dChunks.release(); s.hTiles.release(); s.dTiles.release(); s.dMag.release();
  }
  for (OpRt &o : pl->ops) { if (o.dSharpW) cudaFree(o.dSharpW); if (o.dTw) cudaFree(o.dTw); o.dRaw.release(); if (o.dPitchTab) cudaFree(o.dPitchTab); if (o.dFmtD) cudaFree(o.dFmtD); if (o.dCosTab) cudaFree(o.dCosTab); o.dShs.release(); o.dLag.release(); }
  if (pl->dErr) cudaFree(pl->dErr);
  if (pl->auxStream) cudaStreamDestroy(pl->auxStream);
  if (pl->evFork) cudaEventDestroy(pl->evFork);
  if (pl->evJoin) cudaEventDestroy(pl->evJoin);
  pl->hMeta.release(); pl->dMeta.release(); pl->hPost.release(); pl->dPost.release(); pl->dStat.release(); pl->dMeans.release();
  pl->dPcm.release(); pl->dPcmF.release(); pl->dOut.release();
  if (pl->evMetaDone) cudaEventDestroy(pl->evMetaDone);
  if (pl->evK0) cudaEventDestroy(pl->evK0);
  if (pl->evK1) cudaEventDestroy(pl->evK1);
  if (pl->evKm) cudaEventDestroy(pl->evKm);
  delete pl;
}

int32_t osm_b200_plan_num_elements(const osm_b200_plan *pl) { return pl ? pl->d.nOut : 0; }

const char *osm_b200_plan_element_name(const osm_b200_plan *pl, int32_t idx)
{
  if (!pl || idx < 0 || idx >= (int)pl->d.names.size()) return nullptr;
  return pl->d.names[idx].c_str();
}

double osm_b200_plan_frame_period(const osm_b200_plan *pl) { return pl ? pl->d.fe0().frameStepSec : 0.0; }
int32_t osm_b200_plan_frame_size_samples(const osm_b200_plan *pl) { return pl ? pl->d.fe0().frameSize : 0; }
int32_t osm_b200_plan_frame_step_samples(const osm_b200_plan *pl) { return pl ? pl->d.fe0().frameStep : 0; }
int32_t osm_b200_plan_fft_size(const osm_b200_plan *pl) { return pl ? pl->d.fe0().nfft : 0; }

int64_t osm_b200_plan_num_frames(const osm_b200_plan *pl, int64_t n) { return pl ? desc_num_frames(pl->d, n) : 0; }

int64_t osm_b200_plan_num_time_frames(const osm_b200_plan *pl, int64_t n)
{
  if (!pl || pl->d.groups.empty()) return 0;
  const OutGroup &g = pl->d.groups[0];
  int64_t t = desc_num_static_frames(pl->d, g.stream, n);
  for (int ls : g.limitStreams) t = std::min<int64_t>(t, desc_num_static_frames(pl->d, ls, n));
  return t < 0 ? 0 : t;
}

osm_b200_status osm_b200_plan_frame_offsets(const osm_b200_plan *pl, const int64_t *utt_offsets, int32_t n_utt,
                                            int64_t *frame_offsets)
{
  if (!pl || !utt_offsets || !frame_offsets || n_utt < 0) return fail(OSM_B200_ERR_INVALID, "null argument");
  frame_offsets[0] = 0;
  for (int u = 0; u < n_utt; u++) {
    if (utt_offsets[u + 1] < utt_offsets[u]) return fail(OSM_B200_ERR_INVALID, "utt_offsets must be non-decreasing");
    frame_offsets[u + 1] = frame_offsets[u] + desc_num_frames(pl->d, utt_offsets[u + 1] - utt_offsets[u]);
  }
  return OSM_B200_OK;
}

// (re)build the per-batch tables when the utterance layout changed
static osm_b200_status prepare_batch(osm_b200_plan *pl, const int64_t *uttOff, int nUtt, const int64_t *frameOff,
                                     cudaStream_t st)
{
  const PlanDesc &d = pl->d;
  const bool same = (int)pl->cachedUttOff.size() == nUtt + 1 &&
                    memcmp(pl->cachedUttOff.data(), uttOff, sizeof(int64_t) * (nUtt + 1)) == 0;
  if (!same) {
    // the host tables below are rewritten in place: until the rebuild has completed no layout counts as cached (a rebuild that
    // fails half way must not be mistaken for the previous one)
    pl->cachedUttOff.clear();
    if (pl->metaPending) { CU(cudaEventSynchronize(pl->evMetaDone)); pl->metaPending = false; }
    const size_t nm = (size_t)(nUtt + 1);
    CU(pl->hMeta.reserve(3 * nm));
    long long *hU = pl->hMeta.p, *hR = hU + nm, *hS = hR + nm;
    const int PR = pl->pp.rows;
    const int KT = lld_max_chunk_tiles();
    const bool needPost = pl->pp.nGroups > 0 && !pl->fused;
    size_t nPost = 0;
    pl->uttPost0.assign(nm, 0);
    hR[0] = 0; hS[0] = 0;
    for (int u = 0; u < nUtt; u++) {
      if (uttOff[u + 1] < uttOff[u]) return fail(OSM_B200_ERR_INVALID, "utt_offsets must be non-decreasing");
      const int64_t L = uttOff[u + 1] - uttOff[u];
      hU[u] = uttOff[u];
      hR[u + 1] = hR[u] + desc_num_frames(d, L);
      hS[u + 1] = hS[u] + desc_max_static_frames(d, L);
      pl->uttPost0[u] = (int32_t)nPost;
      if (needPost) nPost += (size_t)((hR[u + 1] - hR[u] + PR - 1) / PR);
    }
    hU[nUtt] = uttOff[nUtt];
    pl->uttPost0[nUtt] = (int32_t)nPost;
    CU(pl->hPost.reserve(nPost + 1));
    {
      size_t ti = 0;
      if (needPost)
        for (int u = 0; u < nUtt; u++) {
          const int64_t To = hR[u + 1] - hR[u];
          for (int64_t r0 = 0; r0 < To; r0 += PR) pl->hPost.p[ti++] = TileRef{u, (int32_t)r0};
        }
    }
    pl->nPostTiles = nPost;
    pl->totalWork = 0;
    // per stream: chunks for lld_kernel, tiles for the standalone ops
    for (size_t si = 0; si < pl->st.size(); si++) {
      StreamRt &rt = pl->st[si];
      const int F = rt.tileF, H = rt.kp.halo;
      std::vector<ChunkRef> chunks;
      std::vector<OpTile> tiles;
      rt.uttChunk0.assign(nm, 0); rt.uttTile0.assign(nm, 0);
      int64_t tileCount = 0;
      for (int u = 0; u < nUtt; u++) {
        const int64_t L = uttOff[u + 1] - uttOff[u];
        const int64_t T = desc_num_static_frames(d, (int)si, L);
        rt.uttChunk0[u] = (int32_t)chunks.size();
        rt.uttTile0[u] = (int32_t)tileCount;
        // chunks: output rows [a,b) whose static range [a-H, b+H) /\ [0,T) is a whole number of
        // tiles (except at the utterance end), at most KT tiles each
        for (int64_t a = 0; a < T;) {
          const int64_t s0 = std::max<int64_t>(a - H, 0);
          const int64_t maxEnd = s0 + (int64_t)F * KT;
          const int64_t b2 = (maxEnd >= T) ? T : maxEnd - H;
          chunks.push_back(ChunkRef{u, (int32_t)a, (int32_t)b2, (int32_t)(tileCount + s0 / F)});
          a = b2;
        }
        if (rt.needTiles)
          for (int64_t f0 = 0; f0 < T; f0 += F)
            tiles.push_back(OpTile{u, (int32_t)f0, (int32_t)std::min<int64_t>(F, T - f0), f0 > 0 ? 1 : 0});
        tileCount += (T + F - 1) / F;
      }
      rt.uttChunk0[nUtt] = (int32_t)chunks.size();
      rt.uttTile0[nUtt] = (int32_t)tileCount;
      rt.nChunks = rt.runLld ? chunks.size() : 0;
      rt.nTiles = tiles.size();
      pl->totalWork += rt.nChunks + rt.nTiles;
      if (rt.runLld) {
        CU(rt.hChunks.reserve(chunks.size() + 1));
        if (!chunks.empty()) memcpy(rt.hChunks.p, chunks.data(), chunks.size() * sizeof(ChunkRef));
        CU(rt.dChunks.reserve(chunks.size() + 1));
        if (!chunks.empty()) CU(cudaMemcpyAsync(rt.dChunks.p, rt.hChunks.p, chunks.size() * sizeof(ChunkRef), cudaMemcpyHostToDevice, st));
        if (rt.kp.opKind < 0 || d.streams[si].dumpMag)
          CU(rt.dMag.reserve((size_t)tileCount * d.streams[si].fe.nBins * F + 64));
      }
      if (rt.needTiles) {
        CU(rt.hTiles.reserve(tiles.size() + 1));
        if (!tiles.empty()) memcpy(rt.hTiles.p, tiles.data(), tiles.size() * sizeof(OpTile));
        CU(rt.dTiles.reserve(tiles.size() + 1));
        if (!tiles.empty()) CU(cudaMemcpyAsync(rt.dTiles.p, rt.hTiles.p, tiles.size() * sizeof(OpTile), cudaMemcpyHostToDevice, st));
      }
    }
    pl->totalRows = hR[nUtt];
    pl->totalStat = hS[nUtt];
    pl->totalSamples = uttOff[nUtt];
    CU(pl->dMeta.reserve(3 * nm));
    CU(pl->dPost.reserve(nPost + 1));
    CU(cudaMemcpyAsync(pl->dMeta.p, pl->hMeta.p, 3 * nm * sizeof(long long), cudaMemcpyHostToDevice, st));
    if (nPost) CU(cudaMemcpyAsync(pl->dPost.p, pl->hPost.p, nPost * sizeof(TileRef), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(pl->evMetaDone, st));
    pl->metaPending = true;
    pl->cachedUttOff.assign(uttOff, uttOff + nUtt + 1);
  }
  if (frameOff) {
    // the caller's row offsets must be the ones the plan derives (bit-exact integer rule)
    const size_t nm = (size_t)(nUtt + 1);
    const long long *hR = pl->hMeta.p + nm;
    for (int u = 0; u <= nUtt; u++)
      if ((long long)frameOff[u] != hR[u]) return fail(OSM_B200_ERR_INVALID, "frame_offsets do not match osm_b200_plan_frame_offsets()");
  }
  return OSM_B200_OK;
}

static cudaError_t prof_mark(osm_b200_plan *pl, const char *name, cudaStream_t st)
{
  if (!pl->profile) return cudaSuccess;
  if (pl->profN >= (int)pl->profEv.size()) {
    cudaEvent_t e;
    cudaError_t r = cudaEventCreate(&e);
    if (r != cudaSuccess) return r;
    pl->profEv.push_back(e);
    pl->profName.push_back(name);
  }
  pl->profName[pl->profN] = name;
  return cudaEventRecord(pl->profEv[pl->profN++], st);
}
#define PROF(name) CU(prof_mark(pl, name, st))

// launch the kernels for utterances [u0, u1) of the prepared batch
static osm_b200_status launch_range(osm_b200_plan *pl, const void *d_pcm, float *d_out, int n_utt, int u0, int u1,
                                    cudaStream_t st)
{
  const PlanDesc &d = pl->d;
  const size_t nm = (size_t)(n_utt + 1);
  const long long *dU = pl->dMeta.p, *dR = dU + nm, *dS = dR + nm;
  PROF("begin");
  const bool f32in = d.fe0().format != OSM_B200_PCM_S16;
  if (f32in) {
    // inputs in another format than 16-bit integer: one conversion pass into mono floats (the caller reserved dPcmF)
    const long long *hU = pl->hMeta.p;
    const long long f0 = hU[u0], f1 = hU[u1];
    const int nc = d.fe0().nChan;
    if (f1 > f0) {
      pcm_convert_kernel<<<(unsigned)((f1 - f0 + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const unsigned char *>(d_pcm) + (size_t)f0 * nc * sample_bytes(d.fe0().format), d.fe0().format, nc, f1 - f0, pl->dPcmF.p + f0);
      CU(cudaGetLastError());
      pl->lastLaunches++;
    }
    d_pcm = pl->dPcmF.p;
    PROF("pcm_convert_kernel");
  }
  // 1. per stream: FFT front end (+ fused band op / magnitude dump)
  for (size_t si = 0; si < pl->st.size(); si++) {
    StreamRt &rt = pl->st[si];
    if (!rt.runLld) continue;
    const int c0 = rt.uttChunk0[u0], c1 = rt.uttChunk0[u1];
    if (c1 <= c0) continue;
    const long long *hS = pl->hMeta.p + 2 * nm;
    const long long row0 = hS[u0], row1 = hS[u1];
    for (size_t j = 0; j <= rt.extra.size(); j++) {
      PassRt &pr = j == 0 ? rt.pass0x : rt.extra[j - 1];
      LldParams kp = j == 0 ? rt.kp : pr.kp;
      kp.pcm = reinterpret_cast<const int16_t *>(d_pcm);
      kp.uttOff = dU;
      kp.chunks = rt.dChunks.p + c0;
      kp.nChunks = c1 - c0;
      kp.magOut = (j == 0 && d.streams[si].dumpMag) ? rt.dMag.p : nullptr;
      if (pl->staticDirect) {
        kp.out = d_out; kp.outStride = d.nOut; kp.outCol = pl->identityOutCol; kp.rowOff = dR;
      } else {
        kp.out = pl->dStat.p; kp.outStride = d.nStatic; kp.rowOff = dS;
        kp.outCol = pr.op >= 0 ? d.ops[pr.op].outCol : 0;
      }
      if (pr.rasta) {
        CU(pr.dBand.reserve((size_t)pl->totalStat * kp.nBands + 64));
        kp.out = pr.dBand.p; kp.outStride = kp.nBands; kp.outCol = 0; kp.rowOff = dS;
      }
      CU((f32in ? launch_lld_f32 : launch_lld)(kp, d.streams[si].fe.nfft, pl->numSMs, st, &pl->lastInfo));
      pl->lastLaunches++;
      PROF("lld_kernel");
      if (pr.rasta) {
        RastaParams rp = pr.rp;
        rp.band = pr.dBand.p; rp.uttOff = dU; rp.statOff = dS;
        CU(launch_rasta(rp, u0, u1, st));
        PROF("rasta_kernel");
        CU(launch_plp_tail(pr.tail, pr.dBand.p, pl->dStat.p, d.nStatic, d.ops[pr.op].outCol, row0, row1, st));
        pl->lastLaunches += 2;
        PROF("plp_tail_kernel");
      }
    }
  }
  if (u0 == 0) CU(cudaEventRecord(pl->evKm, st));
  // 2. standalone ops
  bool forked = false;
  for (OpRt &o : pl->ops) {
    StreamRt &rt = pl->st[o.stream];
    if ((o.kind == SOP_PITCH || o.kind == SOP_JITTER) && pl->auxStream && !forked && !pl->profile) {
      // everything the chain reads (magnitude level, selector energy) has been launched on `st` by now
      CU(cudaEventRecord(pl->evFork, st));
      CU(cudaStreamWaitEvent(pl->auxStream, pl->evFork, 0));
      forked = true;
    }
    if (o.kind == SOP_HARMONICS) continue;              // reads the pitch and formant columns: launched after the join below
    cudaStream_t ks = (forked && (o.kind == SOP_PITCH || o.kind == SOP_JITTER)) ? pl->auxStream : st;
    if (o.kind == SOP_VECOP) {
      const long long *hS = pl->hMeta.p + 2 * nm;
      CU(launch_vecop_ll1(pl->dStat.p, d.nStatic, o.vSrcCol, o.vN, o.vOutCol, hS[u0], hS[u1], st));
      pl->lastLaunches++;
      PROF("vecop_kernel");
      continue;
    }
    const int t0 = rt.uttTile0[u0], t1 = rt.uttTile0[u1];
    if (t1 <= t0) continue;
    if (o.kind == SOP_MAG) {
      CU(launch_mag_rows(rt.dMag.p + (size_t)t0 * o.vN * rt.tileF, rt.dTiles.p + t0, t1 - t0, rt.tileF, o.vN, dS,
                         pl->dStat.p, d.nStatic, o.vOutCol, st, o.magMode, o.magN, o.magDbNorm, o.magMinDb));
      pl->lastLaunches++;
      PROF("mag_rows_kernel");
      continue;
    }
    if (o.kind == SOP_SPECTRAL) {
      SpectralParams sp = o.sp;
      sp.mag = rt.dMag.p + (size_t)t0 * sp.nSrc * sp.F;
      sp.tiles = rt.dTiles.p + t0; sp.nTiles = t1 - t0;
      sp.statOff = dS; sp.stat = pl->dStat.p;
      CU(launch_spectral(sp, st));
      PROF("spectral_kernel");
    } else if (o.kind == SOP_PITCH) {
      ShsParams sh = o.shs;
      CU(o.dShs.reserve((size_t)pl->totalStat * sh.nShsCols + 64));
      CU(o.dLag.reserve((size_t)n_utt + 64));
      sh.mag = rt.dMag.p + (size_t)t0 * sh.nMag * sh.F;
      sh.tiles = rt.dTiles.p + t0; sh.nTiles = t1 - t0;
      sh.statOff = dS; sh.shs = o.dShs.p;
      CU(launch_shs(sh, ks));
      PROF("shs_kernel");
      ViterbiParams vp = o.vit;
      vp.shs = o.dShs.p; vp.uttOff = dU; vp.statOff = dS; vp.stat = pl->dStat.p; vp.lag = o.dLag.p;
      CU(launch_viterbi(vp, u0, u1, ks));
      pl->lastLaunches++;
      PROF("viterbi_kernel");
    } else if (o.kind == SOP_JITTER) {
      JitterParams jp = o.jit;
      jp.pcm = reinterpret_cast<const int16_t *>(d_pcm);
      jp.uttOff = dU; jp.statOff = dS; jp.stat = pl->dStat.p; jp.errFlag = pl->dErr;
      CU((f32in ? launch_jitter_f32 : launch_jitter)(jp, u0, u1, ks));
      PROF("jitter_kernel");
    } else if (o.kind == SOP_PITCHACF) {
      AcfPitchParams ap = o.ap;
      CU(o.dRaw.reserve((size_t)pl->totalStat + 64));
      ap.mag = rt.dMag.p + (size_t)t0 * ap.nSrc * ap.F;
      ap.tiles = rt.dTiles.p + t0; ap.nTiles = t1 - t0;
      ap.statOff = dS; ap.uttOff = dU; ap.raw = o.dRaw.p; ap.stat = pl->dStat.p; ap.nUtt = n_utt;
      CU(launch_acf_pitch(ap, st));
      PROF("acf_pitch_kernel");
      CU(launch_pitch_smooth(ap, u0, u1, st));
      pl->lastLaunches++;
      PROF("pitch_smooth_kernel");
    } else {
      TimeOpParams tp = o.tp;
      tp.pcm = reinterpret_cast<const int16_t *>(d_pcm);
      tp.uttOff = dU; tp.statOff = dS;
      tp.tiles = rt.dTiles.p + t0; tp.nTiles = t1 - t0;
      tp.stat = pl->dStat.p;
      if (o.kind == SOP_FORMANT) { FormantParams fp = o.fmt; fp.tp = tp; CU((f32in ? launch_formant_f32 : launch_formant)(fp, st)); PROF("formant_kernel"); }
      else {
        CU(o.kind == SOP_ENERGY ? (f32in ? launch_energy_f32 : launch_energy)(tp, st)
                                : (o.kind == SOP_INTENSITY ? (f32in ? launch_intensity_f32 : launch_intensity)(tp, st) : (f32in ? launch_mzcr_f32 : launch_mzcr)(tp, st)));
        PROF(o.kind == SOP_ENERGY ? "energy_kernel" : (o.kind == SOP_INTENSITY ? "intensity_kernel" : "mzcr_kernel"));
      }
    }
    pl->lastLaunches++;
  }
  if (forked) {
    CU(cudaEventRecord(pl->evJoin, pl->auxStream));
    CU(cudaStreamWaitEvent(st, pl->evJoin, 0));
  }
  for (OpRt &o : pl->ops) {
    if (o.kind != SOP_HARMONICS) continue;
    StreamRt &rt = pl->st[o.stream];
    const int t0 = rt.uttTile0[u0], t1 = rt.uttTile0[u1];
    if (t1 <= t0) continue;
    HarmonicsParams hp = o.hrm;
    hp.mag = rt.dMag.p + (size_t)t0 * hp.nb * hp.F;
    hp.tiles = rt.dTiles.p + t0; hp.nTiles = t1 - t0;
    hp.statOff = dS; hp.stat = pl->dStat.p;
    CU(launch_harmonics(hp, st));
    pl->lastLaunches++;
    PROF("harmonics_kernel");
  }
  // 3. temporal stages + assembly of the output rows
  if (pl->pp.nGroups > 0 && !pl->fused) {
    PostParams pp = pl->pp;
    if (pl->staticDirect) { pp.stat = d_out + pl->identityOutCol; pp.statStride = d.nOut; pp.statOff = dR; }
    else { pp.stat = pl->dStat.p; pp.statStride = d.nStatic; pp.statOff = dS; }
    pp.out = d_out; pp.outStride = d.nOut; pp.rowOff = dR; pp.uttOff = dU; pp.nUtt = n_utt;
    pp.tiles = pl->dPost.p + pl->uttPost0[u0]; pp.nTiles = pl->uttPost0[u1] - pl->uttPost0[u0];
    pp.means = nullptr;
    if (pl->needMeans) {
      CU(pl->dMeans.reserve((size_t)n_utt * d.nStatic + 64));
      CU(launch_cms_means(pp, pl->dMeans.p, u0, u1, st));
      pl->lastLaunches++;
      PROF("cms_means_kernel");
      pp.means = pl->dMeans.p;
    }
    if (pp.nTiles > 0) {
      CU(launch_post(pp, st));
      pl->lastLaunches++;
      PROF("post_kernel");
    }
  }
  // 4. levels behind the Viterbi-smoothed pitch chain
  if (pl->sp.nGroups > 0 && pl->seqLagOp >= 0) {
    SeqPostParams sq = pl->sp;
    sq.stat = pl->dStat.p; sq.statStride = d.nStatic; sq.statOff = dS; sq.rowOff = dR; sq.uttOff = dU;
    sq.out = d_out; sq.outStride = d.nOut; sq.lag = pl->ops[pl->seqLagOp].dLag.p;
    CU(launch_seq_post(sq, u0, u1, st));
    pl->lastLaunches++;
    PROF("seq_post_kernel");
  }
  return OSM_B200_OK;
}

osm_b200_status osm_b200_plan_run_device(osm_b200_plan *pl, const void *d_pcm, const int64_t *utt_offsets,
                                         int32_t n_utt, const int64_t *frame_offsets, float *d_out, void *stream)
{
  if (!pl || !utt_offsets || n_utt < 0) return fail(OSM_B200_ERR_INVALID, "null argument");
  if (pl->device < 0) return fail(OSM_B200_ERR_CUDA, "description-only plan (device < 0) cannot run; no CPU fallback");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(pl->device));
  pl->lastLaunches = 0;
  pl->timed = false;
  pl->profN = 0;
  osm_b200_status s = prepare_batch(pl, utt_offsets, n_utt, frame_offsets, st);
  if (s != OSM_B200_OK) return s;
  if (pl->totalRows == 0 || pl->totalWork == 0) return OSM_B200_OK;
  if (!d_pcm || !d_out) return fail(OSM_B200_ERR_INVALID, "null device buffer");
  if (pl->d.fe0().format != OSM_B200_PCM_S16) CU(pl->dPcmF.reserve((size_t)utt_offsets[n_utt] + 16));
  if (!pl->staticDirect) CU(pl->dStat.reserve((size_t)pl->totalStat * pl->d.nStatic + 64));
  CU(cudaEventRecord(pl->evK0, st));
  s = launch_range(pl, d_pcm, d_out, n_utt, 0, n_utt, st);
  if (s != OSM_B200_OK) return s;
  CU(cudaEventRecord(pl->evK1, st));
  pl->timed = true;
  return OSM_B200_OK;
}

// Host buffers.  The batch is cut into pieces of whole utterances and pipelined over three
// streams: H2D of piece k+1, kernels of piece k and D2H of piece k-1 overlap (the two copy
// directions use separate copy engines).  Host memory should be pinned (cudaHostAlloc /
// torch pin_memory) for the copies to be asynchronous; pageable memory works but serialises.
static osm_b200_status run_host_impl(osm_b200_plan *pl, const void *pcm, const int64_t *utt_offsets, int32_t n_utt,
                                     const int64_t *frame_offsets, float *out, bool resident)
{
  if (!pl || !utt_offsets || n_utt < 0) return fail(OSM_B200_ERR_INVALID, "null argument");
  if (pl->device < 0) return fail(OSM_B200_ERR_CUDA, "description-only plan (device < 0) cannot run; no CPU fallback");
  CU(cudaSetDevice(pl->device));
  if (!pl->hostStream) {
    CU(cudaStreamCreateWithFlags(&pl->hostStream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&pl->h2dStream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&pl->d2hStream, cudaStreamNonBlocking));
  }
  cudaStream_t st = pl->hostStream;
  pl->lastLaunches = 0;
  pl->timed = false;
  osm_b200_status s = prepare_batch(pl, utt_offsets, n_utt, frame_offsets, st);
  if (s != OSM_B200_OK) return s;
  const long long rows = pl->totalRows;
  if (rows == 0 || pl->totalWork == 0) return OSM_B200_OK;
  if (!pcm || (!out && !resident)) return fail(OSM_B200_ERR_INVALID, "null host buffer");
  const int nOut = pl->d.nOut;
  const int64_t frameBytes = (int64_t)pl->d.fe0().nChan * sample_bytes(pl->d.fe0().format);
  const int64_t totalBytes = utt_offsets[n_utt] * frameBytes;
  CU(pl->dPcm.reserve((size_t)(totalBytes + 1) / 2 + 16));
  if (pl->d.fe0().format != OSM_B200_PCM_S16) CU(pl->dPcmF.reserve((size_t)utt_offsets[n_utt] + 16));
  CU(pl->dOut.reserve((size_t)rows * nOut));
  if (!pl->staticDirect) CU(pl->dStat.reserve((size_t)pl->totalStat * pl->d.nStatic + 64));
  const long long *hR = pl->hMeta.p + (size_t)(n_utt + 1);

  // pieces of ~24 MB of PCM, at most 16, cut at utterance boundaries
  // (plans with per-utterance sequential kernels -- Viterbi, jitter -- get their parallelism from the number of
  // utterances in flight: fewer, larger pieces)
  const int64_t maxPieces = pl->sp.nGroups > 0 ? 4 : 16;
  int nPieces = (int)std::min<int64_t>(maxPieces, std::max<int64_t>(1, totalBytes / (24 << 20)));
  if (nPieces > n_utt) nPieces = n_utt;
  while ((int)pl->evPiece.size() < 2 * nPieces) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    pl->evPiece.push_back(e);
  }
  CU(cudaEventRecord(pl->evK0, st));
  int u0 = 0;
  for (int k = 0; k < nPieces; k++) {
    int u1 = u0;
    const int64_t target = utt_offsets[n_utt] * (int64_t)(k + 1) / nPieces;
    while (u1 < n_utt && (utt_offsets[u1 + 1] <= target || u1 == u0)) u1++;
    if (k == nPieces - 1) u1 = n_utt;
    if (u1 == u0) continue;
    const int64_t sa = utt_offsets[u0] * frameBytes, sb = utt_offsets[u1] * frameBytes;
    CU(cudaMemcpyAsync(reinterpret_cast<unsigned char *>(pl->dPcm.p) + sa, reinterpret_cast<const unsigned char *>(pcm) + sa, (size_t)(sb - sa),
                       cudaMemcpyHostToDevice, pl->h2dStream));
    CU(cudaEventRecord(pl->evPiece[2 * k], pl->h2dStream));
    CU(cudaStreamWaitEvent(st, pl->evPiece[2 * k], 0));
    s = launch_range(pl, pl->dPcm.p, pl->dOut.p, n_utt, u0, u1, st);
    if (s != OSM_B200_OK) {
      // copies of earlier pieces into the caller's buffers may still be in flight: let them land before reporting the error
      cudaStreamSynchronize(pl->h2dStream); cudaStreamSynchronize(st); cudaStreamSynchronize(pl->d2hStream);
      return s;
    }
    CU(cudaEventRecord(pl->evPiece[2 * k + 1], st));
    CU(cudaStreamWaitEvent(pl->d2hStream, pl->evPiece[2 * k + 1], 0));
    const long long ra = hR[u0], rb = hR[u1];
    if (rb > ra && out)
      CU(cudaMemcpyAsync(out + ra * nOut, pl->dOut.p + ra * nOut, (size_t)(rb - ra) * nOut * sizeof(float),
                         cudaMemcpyDeviceToHost, pl->d2hStream));
    u0 = u1;
  }
  CU(cudaEventRecord(pl->evK1, st));
  pl->timed = true;
  CU(cudaStreamSynchronize(pl->d2hStream));
  CU(cudaStreamSynchronize(st));
  if (pl->dErr && pl->sp.nGroups + (int)pl->ops.size() > 0) {
    int flag = 0;
    CU(cudaMemcpy(&flag, pl->dErr, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) {
      CU(cudaMemset(pl->dErr, 0, sizeof(int)));
      return fail(OSM_B200_ERR_UNSUPPORTED, "cPitchJitter: a frame left the supported geometry (F0 period / read window too long for the kernel's workspace, or a read past the end of the utterance); its rows were zeroed");
    }
  }
  return OSM_B200_OK;
}

osm_b200_status osm_b200_plan_run_host(osm_b200_plan *pl, const void *pcm, const int64_t *utt_offsets, int32_t n_utt,
                                       const int64_t *frame_offsets, float *out)
{
  return run_host_impl(pl, pcm, utt_offsets, n_utt, frame_offsets, out, false);
}

// Host PCM in, rows left in HBM: the H2D pipeline and kernels of run_host without the copy back.  *d_rows = the plan's own
// row buffer ([rows][num_elements], valid until the plan's next run): the hand-over point to a consumer that summarises the
// rows on the device (osm_b200_functionals_run_device).
osm_b200_status osm_b200_plan_run_host_resident(osm_b200_plan *pl, const void *pcm, const int64_t *utt_offsets, int32_t n_utt,
                                                const int64_t *frame_offsets, const float **d_rows)
{
  if (!d_rows) return fail(OSM_B200_ERR_INVALID, "null argument");
  *d_rows = nullptr;
  osm_b200_status s = run_host_impl(pl, pcm, utt_offsets, n_utt, frame_offsets, nullptr, true);
  if (s == OSM_B200_OK) *d_rows = pl->dOut.p;
  return s;
}

osm_b200_status osm_b200_window_table(const osm_b200_windower *w, int32_t n, float *out)
{
  if (!w || !out || n <= 0) return fail(OSM_B200_ERR_INVALID, "null argument");
  if (w->winFunc < OSM_B200_WIN_RECTANGLE || w->winFunc > OSM_B200_WIN_LANCZOS) return fail(OSM_B200_ERR_INVALID, "unknown window function");
  const double al[4] = {w->alpha0, w->alpha1, w->alpha2, w->alpha3};
  std::vector<float> t;
  build_window(w->winFunc, n, w->sigma, w->gain, t, al, w->squareRoot, std::min(std::max(w->fade, 0.0), 0.5));
  memcpy(out, t.data(), sizeof(float) * (size_t)n);
  return OSM_B200_OK;
}

int64_t osm_b200_plan_num_frames_first_eoi(const osm_b200_plan *pl, int64_t n) { return pl ? desc_num_frames_first_eoi(pl->d, n) : 0; }
int64_t osm_b200_plan_num_frames_first_eoi_v(const osm_b200_plan *pl, int64_t n, int64_t v) { return pl ? desc_num_frames_first_eoi(pl->d, n, v) : 0; }

osm_b200_status osm_b200_plan_copy_seq_lag(osm_b200_plan *pl, int32_t *out, int32_t n_utt)
{
  if (!pl || !out || n_utt < 0) return fail(OSM_B200_ERR_INVALID, "null argument");
  for (int u = 0; u < n_utt; u++) out[u] = -1;
  if (pl->device < 0 || pl->seqLagOp < 0 || n_utt == 0) return OSM_B200_OK;
  CU(cudaSetDevice(pl->device));
  CU(cudaDeviceSynchronize());
  if (!pl->ops[pl->seqLagOp].dLag.p) return OSM_B200_OK;
  CU(cudaMemcpy(out, pl->ops[pl->seqLagOp].dLag.p, sizeof(int) * (size_t)n_utt, cudaMemcpyDeviceToHost));
  return OSM_B200_OK;
}

int32_t osm_b200_plan_last_launch_count(const osm_b200_plan *pl) { return pl ? pl->lastLaunches : 0; }

int32_t osm_b200_plan_sample_frame_bytes(const osm_b200_plan *pl) { return pl ? pl->d.fe0().nChan * sample_bytes(pl->d.fe0().format) : 0; }

int32_t osm_b200_plan_take_device_flags(osm_b200_plan *pl)
{
  if (!pl || pl->device < 0 || !pl->dErr) return 0;
  int flag = 0;
  if (cudaSetDevice(pl->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(&flag, pl->dErr, sizeof flag, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (flag && cudaMemset(pl->dErr, 0, sizeof(int)) != cudaSuccess) return -1;
  return flag;
}

float osm_b200_plan_last_kernel_ms(osm_b200_plan *pl)
{
  if (!pl || !pl->timed) return -1.f;
  if (cudaEventSynchronize(pl->evK1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, pl->evK0, pl->evK1) != cudaSuccess) return -1.f;
  return ms;
}

void osm_b200_plan_set_profiling(osm_b200_plan *pl, int32_t on) { if (pl) pl->profile = on != 0; }

int32_t osm_b200_plan_profile_count(osm_b200_plan *pl) { return (pl && pl->profN > 1) ? pl->profN - 1 : 0; }

osm_b200_status osm_b200_plan_profile_entry(osm_b200_plan *pl, int32_t idx, const char **name, float *ms)
{
  if (!pl || idx < 0 || idx + 1 >= pl->profN) return fail(OSM_B200_ERR_INVALID, "profile entry out of range");
  CU(cudaSetDevice(pl->device));
  CU(cudaEventSynchronize(pl->profEv[idx + 1]));
  float t = 0.0f;
  CU(cudaEventElapsedTime(&t, pl->profEv[idx], pl->profEv[idx + 1]));
  if (name) *name = pl->profName[idx + 1];
  if (ms) *ms = t;
  return OSM_B200_OK;
}

osm_b200_status osm_b200_plan_last_kernel_times(osm_b200_plan *pl, float *lld_ms, float *post_ms)
{
  if (!pl || !pl->timed) return fail(OSM_B200_ERR_INVALID, "no timed run");
  CU(cudaEventSynchronize(pl->evK1));
  float a = 0.f, b = 0.f;
  CU(cudaEventElapsedTime(&a, pl->evK0, pl->evKm));
  CU(cudaEventElapsedTime(&b, pl->evKm, pl->evK1));
  if (lld_ms) *lld_ms = a;
  if (post_ms) *post_ms = b;
  return OSM_B200_OK;
}

}  // extern "C"
