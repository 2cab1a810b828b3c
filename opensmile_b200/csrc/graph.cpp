// graph.cpp -- resolves a component list (the .conf graph as handed over the C ABI) into a
// fused plan description.  This is the host-side analogue of cComponentManager's
// configure/finalise phase (src/core/componentManager.cpp:606-838) restricted to the LLD
// sub-graph: it follows reader.dmLevel / writer.dmLevel wiring, applies the per-component
// geometry rules (frame size rounding, FFT size, frameSizeSec rescale, frame counts, field
// names) and emits tables.  Citations relative to /root/reference/src.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

#include "plan.hpp"

namespace osm {

namespace {

struct Resolver {
  const osm_b200_component *comps;
  int n;
  std::map<std::string, int> producer;  // level name -> component index
  std::string err;

  const osm_b200_component *prod(const char *level) const {
    auto it = producer.find(level);
    return it == producer.end() ? nullptr : &comps[it->second];
  }
};

const char *type_name(int t)
{
  static const char *names[] = {"cWaveSource", "cFramer", "cVectorPreemphasis", "cWindower",
    "cTransformFFT", "cFFTmagphase", "cMelspec", "cMfcc", "cPlp", "cSpectral", "cEnergy",
    "cMZcr", "cAcf", "cPitchACF", "cDeltaRegression", "cContourSmoother", "cVectorConcat",
    "cVectorOperation"};
  return (t >= 0 && t < OSM_B200_C_COUNT_) ? names[t] : "?";
}

// default nameAppend per type (ConfigType defaults: dspcore/deltaRegression.cpp:34,
// dspcore/contourSmoother.cpp:33, lldcore/mfcc.cpp:33, dspcore/acf.cpp:46, lldcore/energy.cpp:33)
const char *default_name_append(int t)
{
  switch (t) {
    case OSM_B200_C_MFCC: return "mfcc";
    case OSM_B200_C_DELTAREGRESSION: return "de";
    case OSM_B200_C_CONTOURSMOOTHER: return "sma";
    case OSM_B200_C_ACF: return "acf";
    case OSM_B200_C_ENERGY: return "energy";
    default: return "";
  }
}

// cDataProcessor::addNameAppendFieldAuto (core/dataProcessor.cpp:272-325)
std::string name_append_auto(const osm_b200_component &c, const std::string &base, const char *customFixed)
{
  std::string na = c.nameAppend[0] ? c.nameAppend : default_name_append(c.type);
  std::string tail = std::string(customFixed ? customFixed : "") + na;
  if (!tail.empty()) {
    if (c.copyInputName && !base.empty()) return base + "_" + tail;
    return tail;
  }
  if (c.copyInputName && !base.empty()) return base;
  return "noname";
}

}  // namespace

// core/winToVecProcessor.cpp:868-877 (noPostEOIprocessing=1, frameCenterSpecial=left):
// only complete frames => T = floor((L - size)/step) + 1
int64_t desc_num_static_frames(const PlanDesc &d, int64_t L)
{
  if (L < d.fe.frameSize || d.fe.frameSize <= 0) return 0;
  return (L - d.fe.frameSize) / d.fe.frameStep + 1;
}

// window processors emit T + W frames at EOI (core/dataMemoryLevel.cpp:1022-1026 via
// core/windowProcessor.cpp:85-119); cVectorConcat emits min over its inputs
// (core/dataReader.cpp:375-380).
int64_t desc_num_frames(const PlanDesc &d, int64_t L)
{
  const int64_t T = desc_num_static_frames(d, L);
  if (T <= 0) return 0;
  int64_t best = -1;
  for (const auto &g : d.groups) {
    int64_t t = T;
    for (const auto &s : g.stages) t += s.win;
    if (best < 0 || t < best) best = t;
  }
  return best < 0 ? 0 : best;
}

osm_b200_status compile_graph(const osm_b200_component *comps, int n, const char *outputLevel,
                              PlanDesc &d, std::string &err)
{
  char buf[512];
  Resolver R{comps, n, {}, {}};
  int nWave = 0;
  for (int i = 0; i < n; i++) {
    if (comps[i].type < 0 || comps[i].type >= OSM_B200_C_COUNT_) { err = "unknown component type"; return OSM_B200_ERR_INVALID; }
    if (comps[i].type == OSM_B200_C_WAVESOURCE) nWave++;
    if (!comps[i].writer_dmLevel[0]) { err = "component without writer.dmLevel"; return OSM_B200_ERR_INVALID; }
    if (R.producer.count(comps[i].writer_dmLevel)) {
      snprintf(buf, sizeof buf, "level '%s' has more than one writer", comps[i].writer_dmLevel);
      err = buf; return OSM_B200_ERR_INVALID;   // one writer per level (core/dataWriter.cpp)
    }
    R.producer[comps[i].writer_dmLevel] = i;
  }
  if (nWave != 1) { err = "graph must contain exactly one cWaveSource"; return OSM_B200_ERR_INVALID; }
  if (!outputLevel || !R.prod(outputLevel)) { err = "output level has no writer"; return OSM_B200_ERR_INVALID; }

  // ---- output level: a cVectorConcat of chains, or a single chain ----
  std::vector<std::string> chainLevels;
  const osm_b200_component *outc = R.prod(outputLevel);
  if (outc->type == OSM_B200_C_VECTORCONCAT) {
    for (int i = 0; i < outc->n_inputs; i++) chainLevels.push_back(outc->reader_dmLevel[i]);
  } else {
    chainLevels.push_back(outputLevel);
  }

  auto single_input = [&](const osm_b200_component *c) -> const osm_b200_component * {
    if (c->n_inputs != 1) return nullptr;
    return R.prod(c->reader_dmLevel[0]);
  };

  std::map<const osm_b200_component *, int> staticOpOf;   // static producer -> op index
  const osm_b200_component *feTail = nullptr;              // the cFFTmagphase all ops hang off
  std::vector<std::string> staticBaseName;                 // per op: field base name

  d = PlanDesc();
  for (const std::string &lvl : chainLevels) {
    // walk back through temporal stages
    const osm_b200_component *c = R.prod(lvl.c_str());
    if (!c) { err = "level '" + lvl + "' has no writer"; return OSM_B200_ERR_INVALID; }
    std::vector<const osm_b200_component *> stageComps;
    while (c && (c->type == OSM_B200_C_DELTAREGRESSION || c->type == OSM_B200_C_CONTOURSMOOTHER)) {
      stageComps.insert(stageComps.begin(), c);
      c = single_input(c);
    }
    if (!c) { err = "broken temporal chain below level '" + lvl + "'"; return OSM_B200_ERR_INVALID; }

    // static feature producer
    int opIdx;
    auto it = staticOpOf.find(c);
    if (it != staticOpOf.end()) {
      opIdx = it->second;
    } else {
      if (c->type != OSM_B200_C_MFCC && c->type != OSM_B200_C_PLP) {
        snprintf(buf, sizeof buf, "component '%s' (%s) is not a supported static LLD producer", c->name, type_name(c->type));
        err = buf; return OSM_B200_ERR_UNSUPPORTED;
      }
      const osm_b200_component *mel = single_input(c);
      if (!mel || mel->type != OSM_B200_C_MELSPEC) { err = "cMfcc / cPlp must read a cMelspec level"; return OSM_B200_ERR_UNSUPPORTED; }
      const osm_b200_component *mag = single_input(mel);
      if (!mag || mag->type != OSM_B200_C_FFTMAGPHASE) { err = "cMelspec must read a cFFTmagphase level"; return OSM_B200_ERR_UNSUPPORTED; }
      if (feTail && feTail != mag) { err = "all static LLDs must share one framer/FFT chain"; return OSM_B200_ERR_UNSUPPORTED; }

      if (!feTail) {
        // ---- resolve the front end once: fftmag <- fft <- win <- [pe] <- frame <- wave ----
        feTail = mag;
        const auto &mp = mag->u.fftmagphase;
        if (!mp.magnitude || mp.phase || mp.normalise || mp.power || mp.dBpsd) {
          err = "cFFTmagphase: only magnitude=1 (no phase/normalise/power/dBpsd) is supported"; return OSM_B200_ERR_UNSUPPORTED;
        }
        const osm_b200_component *fft = single_input(mag);
        if (!fft || fft->type != OSM_B200_C_TRANSFORMFFT) { err = "cFFTmagphase must read a cTransformFFT level"; return OSM_B200_ERR_UNSUPPORTED; }
        if (fft->u.transformfft.inverse) { err = "cTransformFFT.inverse=1 is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *win = single_input(fft);
        if (!win || win->type != OSM_B200_C_WINDOWER) { err = "cTransformFFT must read a cWindower level"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *x = single_input(win);
        const osm_b200_component *pe = nullptr;
        if (x && x->type == OSM_B200_C_VECTORPREEMPHASIS) { pe = x; x = single_input(x); }
        if (!x || x->type != OSM_B200_C_FRAMER) { err = "cWindower must read a cFramer (optionally via cVectorPreemphasis)"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *frm = x;
        const osm_b200_component *wav = single_input(frm);
        if (!wav || wav->type != OSM_B200_C_WAVESOURCE) { err = "cFramer must read the cWaveSource level"; return OSM_B200_ERR_UNSUPPORTED; }

        FrontEnd &fe = d.fe;
        const auto &wp = wav->u.wavesource;
        if (wp.sampleRate <= 0 || wp.nChannels < 1) { err = "cWaveSource: bad sampleRate/nChannels"; return OSM_B200_ERR_INVALID; }
        if (wp.nChannels > 1 && !wp.monoMixdown) { err = "multi-channel without monoMixdown is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        if (wp.format != OSM_B200_PCM_S16) { err = "only 16-bit integer PCM is supported"; return OSM_B200_ERR_UNSUPPORTED; }
        fe.sampleRate = wp.sampleRate; fe.nChan = wp.nChannels; fe.format = wp.format; fe.mixdown = true;

        // cWinToVecProcessor::configureWriter (core/winToVecProcessor.cpp:435-456)
        const auto &fp = frm->u.framer;
        if (!fp.frameCenterSpecialLeft) { err = "cFramer: only frameCenterSpecial=left is supported"; return OSM_B200_ERR_UNSUPPORTED; }
        if (!fp.noPostEOIprocessing) { err = "cFramer: only noPostEOIprocessing=1 is supported"; return OSM_B200_ERR_UNSUPPORTED; }
        const double T = 1.0 / wp.sampleRate;
        double frameSize = fp.frameSize, frameStep = fp.frameStep;
        long fsf = (long)round(frameSize / T);
        if (frameStep == 0.0) frameStep = frameSize;
        long fstf = (long)round(frameStep / T);
        if (fstf == 0) fstf = fsf;
        if (fsf < 2) { err = "cFramer: frame too short"; return OSM_B200_ERR_INVALID; }
        fe.frameSize = (int)fsf; fe.frameStep = (int)fstf;
        fe.frameSizeSec = frameSize; fe.frameStepSec = frameStep;

        if (pe) {
          fe.preemph = true;
          fe.preK = (float)pe->u.vectorpreemphasis.k;     // dspcore/vectorPreemphasis.cpp:55
          fe.preDe = pe->u.vectorpreemphasis.de;
        }
        const auto &wnp = win->u.windower;
        build_window(wnp.winFunc, fe.frameSize, wnp.sigma, wnp.gain, fe.window);
        fe.winOffset = (float)wnp.offset;

        // cTransformFFT: next power of two >= frame size, >= 4 (dspcore/transformFft.cpp:124-129);
        // frameSizeSec *= nfft/frameSize (:78-85)
        int nfft = 4;
        while (nfft < fe.frameSize) nfft <<= 1;
        fe.nfft = nfft; fe.nBins = nfft / 2 + 1;
        fe.fftFrameSizeSec = frameSize;
        if (nfft != fe.frameSize) fe.fftFrameSizeSec *= (double)nfft / (double)fe.frameSize;
        fe.zeroPadSymmetric = fft->u.transformfft.zeroPadSymmetric != 0;
        if (nfft < 64 || nfft > 4096) { err = "FFT size out of the supported range 64..4096"; return OSM_B200_ERR_UNSUPPORTED; }
      }

      // field base name along the chain: outFieldName -> (pe/win/fft keep) -> fftMag -> melspec keeps
      const osm_b200_component *fft = single_input(mag);
      const osm_b200_component *win = single_input(fft);
      const osm_b200_component *x = single_input(win);
      if (x->type == OSM_B200_C_VECTORPREEMPHASIS) x = single_input(x);
      const osm_b200_component *wav = single_input(x);
      std::string base = wav->u.wavesource.outFieldName[0] ? wav->u.wavesource.outFieldName : "pcm";
      base = name_append_auto(*mag, base, "fftMag");         // dspcore/fftmagphase.cpp:154
      base = name_append_auto(*mel, base, nullptr);          // melspec keeps the name

      // mel bank + mfcc op
      const auto &melp = mel->u.melspec;
      if (melp.nBands < 1 || melp.nBands > 64 || melp.nBands >= d.fe.nBins) { err = "cMelspec.nBands out of range"; return OSM_B200_ERR_UNSUPPORTED; }
      MelBank mb;
      build_mel(melp, d.fe.nBins, d.fe.fftFrameSizeSec, mb);
      d.mels.push_back(mb);
      StaticOp op;
      std::string opName;
      if (c->type == OSM_B200_C_MFCC) {
        op.kind = SOP_MFCC;
        const auto &mfp = c->u.mfcc;
        if (mfp.lastMfcc < mfp.firstMfcc || mfp.firstMfcc < 0 || mfp.lastMfcc >= melp.nBands) { err = "cMfcc: bad firstMfcc/lastMfcc"; return OSM_B200_ERR_INVALID; }
        build_mfcc(mfp, melp.nBands, op.mfcc);
        op.mfcc.melIdx = (int)d.mels.size() - 1;
        op.nOut = op.mfcc.nMfcc;
        op.arrNameOffset = op.mfcc.first;                              // lldcore/mfcc.cpp:125
        opName = name_append_auto(*c, base, nullptr);                  // lldcore/mfcc.cpp:120-128
      } else {
        op.kind = SOP_PLP;
        if (!build_plp(c->u.plp, d.mels.back(), op.plp, err)) return OSM_B200_ERR_UNSUPPORTED;
        op.plp.melIdx = (int)d.mels.size() - 1;
        op.nOut = op.plp.nOut;
        op.arrNameOffset = 0;
        // lldcore/plp.cpp:232-267 replaces the field name, then cVectorProcessor appends nameAppend
        const char *fixed = op.plp.doLpToCeps ? "PlpCC" : (op.plp.doLP ? "Plpc" : (op.plp.doIDFT ? "audAutoCor" : "audSpec"));
        opName = name_append_auto(*c, fixed, nullptr);
      }
      op.outCol = d.nStatic;
      d.nStatic += op.nOut;
      d.ops.push_back(op);
      opIdx = (int)d.ops.size() - 1;
      staticOpOf[c] = opIdx;
      staticBaseName.push_back(opName);
    }

    // ---- group ----
    OutGroup g;
    g.srcCol = d.ops[opIdx].outCol;
    g.n = d.ops[opIdx].nOut;
    g.outCol = d.nOut;
    std::string nm = staticBaseName[opIdx];
    for (const osm_b200_component *s : stageComps) {
      Stage st;
      if (s->type == OSM_B200_C_DELTAREGRESSION) {
        const auto &p = s->u.deltaregression;
        if (p.absOutput || p.halfWaveRect || p.onlyInSegments || p.relativeDelta) {
          err = "cDeltaRegression: absOutput/halfWaveRect/onlyInSegments/relativeDelta are not supported"; return OSM_B200_ERR_UNSUPPORTED;
        }
        if (p.deltawin < 1 || p.deltawin > 8) { err = "cDeltaRegression.deltawin must be 1..8"; return OSM_B200_ERR_UNSUPPORTED; }
        st = Stage{ST_DELTA, p.deltawin, 0};
      } else {
        const auto &p = s->u.contoursmoother;
        if (p.smaWin < 1 || (p.smaWin & 1) == 0 || p.smaWin > 9) { err = "cContourSmoother.smaWin must be odd, 1..9"; return OSM_B200_ERR_UNSUPPORTED; }
        st = Stage{ST_SMA, (p.smaWin - 1) / 2, p.noZeroSma};
      }
      g.stages.push_back(st);
      nm = name_append_auto(*s, nm, nullptr);
    }
    if (g.stages.size() > 3) { err = "more than 3 chained temporal stages"; return OSM_B200_ERR_UNSUPPORTED; }
    d.nOut += g.n;
    d.groups.push_back(g);
    // element names: name[idx + arrNameOffset] (core/dataMemoryLevel.cpp:1158-1169);
    // cMfcc passes firstMfcc as arrNameOffset (lldcore/mfcc.cpp:125)
    const int off = d.ops[opIdx].arrNameOffset;
    for (int i = 0; i < g.n; i++) {
      snprintf(buf, sizeof buf, "%s[%d]", nm.c_str(), i + off);
      d.names.push_back(buf);
    }
  }
  if (d.ops.empty()) { err = "empty plan"; return OSM_B200_ERR_INVALID; }
  return OSM_B200_OK;
}

}  // namespace osm
