// graph.cpp -- resolves a component list (the .conf graph as handed over the C ABI) into a
// fused plan description.  This is the host-side analogue of cComponentManager's
// configure/finalise phase (src/core/componentManager.cpp:606-838) restricted to the LLD
// sub-graph: it follows reader.dmLevel / writer.dmLevel wiring, applies the per-component
// geometry rules (frame size rounding, FFT size, frameSizeSec rescale, frame counts, field
// names) and emits tables.  Citations relative to /root/reference/src.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
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
    "cVectorOperation", "cFullinputMean", "cIntensity", "cSpecScale", "cPitchShs", "cPitchSmootherViterbi",
    "cValbasedSelector", "cPitchJitter", "cSpecResample", "cLpc", "cFormantLpc", "cDataSelector", "cHarmonics"};
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
int64_t desc_num_static_frames(const PlanDesc &d, int stream, int64_t L)
{
  const FrontEnd &fe = d.streams[stream].fe;
  if (L < fe.frameSize || fe.frameSize <= 0) return 0;
  return (L - fe.frameSize) / fe.frameStep + 1;
}

int64_t desc_max_static_frames(const PlanDesc &d, int64_t L)
{
  int64_t m = 0;
  for (size_t s = 0; s < d.streams.size(); s++) m = std::max<int64_t>(m, desc_num_static_frames(d, (int)s, L));
  return m;
}

// window processors emit T + W frames at EOI (core/dataMemoryLevel.cpp:1022-1026 via
// core/windowProcessor.cpp:85-119); cVectorConcat emits min over its inputs
// (core/dataReader.cpp:375-380).
int64_t desc_num_frames(const PlanDesc &d, int64_t L)
{
  int64_t best = -1;
  for (const auto &g : d.groups) {
    int64_t t = desc_num_static_frames(d, g.stream, L);
    for (int ls : g.limitStreams) t = std::min<int64_t>(t, desc_num_static_frames(d, ls, L));
    if (t <= 0) return 0;
    for (const auto &s : g.stages) t += s.win;
    if (best < 0 || (d.padRows ? t > best : t < best)) best = t;
  }
  return best < 0 ? 0 : best;
}

// Frames of the output level that exist when a full-input reader (cFunctionals with frameMode = full) ticks for the first time
// after end of input was raised.  The reference's window processors run with blocksize 1: stage s has produced
// c0_s = max(c0_{s-1} - W_s, 0) frames before EOI (tick-order model, see post_kernel) and adds exactly one more in the first EOI
// tick before the components behind it tick; the reader then takes what is there (core/dataReader.cpp:560-585) and, having read
// once, never reads again -- so the frames the window processors append later are NOT part of the summary unless the
// configuration sets EOIlevel (the shipped IS09 / IS10 files do not).  Verified against the reference's functionals rows
// (tests/test_functionals_cpu.py).
// Levels behind the SHS pitch chain (lagKind != 0): the Viterbi smoother has written V frames when end of input is raised
// (data dependent, osm_b200_plan_copy_seq_lag); cPitchJitter does not run while end of input is set (lld/pitchJitter.cpp:593), so the
// chain's static level still holds min(V, T) frames in the first EOI tick: a smoother behind it ends at V rows, its delta at V - 2
// (pinned on the reference's ComParE_2016 functionals rows: V = 143 / 193 / 198 -> 143 | 141, 193 | 191, 198 | 196).
int64_t desc_num_frames_first_eoi(const PlanDesc &d, int64_t L, int64_t V)
{
  int64_t best = -1;
  for (const auto &g : d.groups) {
    int64_t t = desc_num_static_frames(d, g.stream, L);
    for (int ls : g.limitStreams) t = std::min<int64_t>(t, desc_num_static_frames(d, ls, L));
    if (t <= 0) return 0;
    if (g.lagKind != 0 && V >= 0) t = std::min<int64_t>(t, V);
    int64_t c0 = t, fin = t;
    for (const auto &s : g.stages) { c0 = std::max<int64_t>(c0 - s.win, 0); fin += s.win; }
    const int64_t avail = g.stages.empty() ? t : std::min<int64_t>(c0 + 1, fin);
    if (best < 0 || avail < best) best = avail;
  }
  return best < 0 ? 0 : best;
}

namespace {

struct ChainInfo {               // resolved front-end chain below a static producer
  const osm_b200_component *wav = nullptr, *frm = nullptr, *pe = nullptr, *win = nullptr, *fft = nullptr, *mag = nullptr;
};

}  // namespace

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

  auto single_input = [&](const osm_b200_component *c) -> const osm_b200_component * {
    if (!c || c->n_inputs != 1) return nullptr;
    return R.prod(c->reader_dmLevel[0]);
  };

  // walk from a time-domain level (framer / pre-emphasis / windower output) down to the wave source
  auto resolve_time_chain = [&](const osm_b200_component *x, ChainInfo &ci) -> bool {
    if (x && x->type == OSM_B200_C_WINDOWER) { ci.win = x; x = single_input(x); }
    if (x && x->type == OSM_B200_C_VECTORPREEMPHASIS) { ci.pe = x; x = single_input(x); }
    if (!x || x->type != OSM_B200_C_FRAMER) { err = "expected cFramer [-> cVectorPreemphasis] [-> cWindower] below this component"; return false; }
    ci.frm = x;
    ci.wav = single_input(x);
    if (!ci.wav || ci.wav->type != OSM_B200_C_WAVESOURCE) { err = "cFramer must read the cWaveSource level"; return false; }
    return true;
  };
  auto resolve_mag_chain = [&](const osm_b200_component *mag, ChainInfo &ci, bool asOutput = false) -> bool {
    if (!mag || mag->type != OSM_B200_C_FFTMAGPHASE) { err = "expected a cFFTmagphase level"; return false; }
    ci.mag = mag;
    const auto &mp = mag->u.fftmagphase;
    if ((!mp.magnitude && !mp.dBpsd) || mp.phase) { err = "cFFTmagphase: only magnitude=1 without phase is supported"; return false; }
    if (!asOutput && (mp.normalise || mp.power || mp.dBpsd)) {
      err = "cFFTmagphase: normalise / power / dBpsd are supported where the level is the output level only (its consumers read the plain magnitude)"; return false;
    }
    ci.fft = single_input(mag);
    if (!ci.fft || ci.fft->type != OSM_B200_C_TRANSFORMFFT) { err = "cFFTmagphase must read a cTransformFFT level"; return false; }
    if (ci.fft->u.transformfft.inverse) { err = "cTransformFFT.inverse=1 is not supported"; return false; }
    const osm_b200_component *w = single_input(ci.fft);
    if (!w || w->type != OSM_B200_C_WINDOWER) { err = "cTransformFFT must read a cWindower level"; return false; }
    return resolve_time_chain(w, ci);
  };

  d = PlanDesc();
  d.padRows = R.prod(outputLevel)->type == OSM_B200_C_VECTORCONCAT && strcmp(R.prod(outputLevel)->name, "_unionconcat") == 0;
  // find or create the stream of a chain; needFft extends an existing time-only stream
  auto get_stream = [&](const ChainInfo &ci, bool needFft, int &idx) -> osm_b200_status {
    for (size_t s = 0; s < d.streams.size(); s++) {
      Stream &st = d.streams[s];
      if (st.keyFramer == ci.frm && st.keyPe == ci.pe && (st.keyWin == ci.win || !ci.win || !st.keyWin)) {
        if (ci.win && !st.keyWin) continue;      // a windowed chain cannot reuse a window-less stream
        if (!ci.win && st.keyWin) { /* framer-level reader on a windowed stream: fine, same geometry */ }
        if (needFft && !st.hasFft) continue;
        idx = (int)s;
        return OSM_B200_OK;
      }
    }
    Stream st;
    FrontEnd &fe = st.fe;
    const auto &wp = ci.wav->u.wavesource;
    if (wp.sampleRate <= 0 || wp.nChannels < 1) { err = "cWaveSource: bad sampleRate/nChannels"; return OSM_B200_ERR_INVALID; }
    if (wp.nChannels > 1 && !wp.monoMixdown) { err = "multi-channel without monoMixdown is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
    if (wp.format < OSM_B200_PCM_S16 || wp.format > OSM_B200_PCM_S32) { err = "cWaveSource: unknown sample format"; return OSM_B200_ERR_INVALID; }
    fe.sampleRate = wp.sampleRate; fe.nChan = wp.nChannels; fe.format = wp.format; fe.mixdown = true;
    // cWinToVecProcessor::configureWriter (core/winToVecProcessor.cpp:435-456)
    const auto &fp = ci.frm->u.framer;
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
    if (ci.pe) {
      fe.preemph = true;
      fe.preK = (float)ci.pe->u.vectorpreemphasis.k;     // dspcore/vectorPreemphasis.cpp:55
      fe.preDe = ci.pe->u.vectorpreemphasis.de;
    }
    if (ci.win) {
      const auto &wnp = ci.win->u.windower;
      { const double al[4] = {wnp.alpha0, wnp.alpha1, wnp.alpha2, wnp.alpha3};
        build_window(wnp.winFunc, fe.frameSize, wnp.sigma, wnp.gain, fe.window, al, wnp.squareRoot, std::min(std::max(wnp.fade, 0.0), 0.5)); }
      fe.winOffset = (float)wnp.offset;
      st.hasWindow = true;
    } else {
      fe.window.assign(fe.frameSize, 1.0f);
    }
    // cTransformFFT: next power of two >= frame size, >= 4 (dspcore/transformFft.cpp:124-129);
    // frameSizeSec *= nfft/frameSize (:78-85)
    int nfft = 4;
    while (nfft < fe.frameSize) nfft <<= 1;
    fe.nfft = nfft; fe.nBins = nfft / 2 + 1;
    fe.fftFrameSizeSec = frameSize;
    if (nfft != fe.frameSize) fe.fftFrameSizeSec *= (double)nfft / (double)fe.frameSize;
    if (ci.fft) fe.zeroPadSymmetric = ci.fft->u.transformfft.zeroPadSymmetric != 0;
    st.hasFft = needFft;
    st.keyFramer = ci.frm; st.keyPe = ci.pe; st.keyWin = ci.win;
    d.streams.push_back(st);
    idx = (int)d.streams.size() - 1;
    return OSM_B200_OK;
  };

  std::map<const osm_b200_component *, int> staticOpOf;   // static producer -> op index

  // ---- leaves of the output level ----
  // The output level is a tree: temporal stages (cDeltaRegression / cContourSmoother, one input
  // level each) and cVectorConcat nodes over static producers.  Temporal stages work element by
  // element, so stage(concat(a, b)) = concat(stage(a), stage(b)) as long as concat does not
  // truncate (all inputs equally long); every leaf becomes one output group carrying the stages
  // between it and the output level.
  struct Leaf { const osm_b200_component *c; std::vector<const osm_b200_component *> stages; bool arraysOnly; };
  std::vector<Leaf> leaves;
  struct ConcatCheck { size_t g0, g1, above; };          // leaves [g0, g1) sit below a concat that has stages above it
  std::vector<std::pair<size_t, size_t>> leafGroups;   // leaf -> its groups [first, last)
  std::vector<ConcatCheck> concatChecks;
  struct SelScope { const osm_b200_component *c; size_t l0, l1, above; };   // leaves [l0, l1) sit below selector c, `above` stages above it
  std::vector<SelScope> selScopes;
  bool inSelector = false;
  // leaves [l0, l1) are the data levels of a cValbasedSelector (zeroVec = 1) that gates them element-wise with a column of the pitch level
  struct GateScope { const osm_b200_component *c; size_t l0, l1, above; };
  std::vector<GateScope> gateScopes;
  std::vector<std::vector<std::string>> leafSelNames;   // per leaf: element names as a selector above it sees them
  std::function<osm_b200_status(const std::string &, std::vector<const osm_b200_component *>, int, bool)> expand =
    [&](const std::string &lvl, std::vector<const osm_b200_component *> above, int depth, bool arraysOnly) -> osm_b200_status {
    if (depth > 8) { err = "level graph nested too deeply (cycle?)"; return OSM_B200_ERR_INVALID; }
    const osm_b200_component *c = R.prod(lvl.c_str());
    if (!c) { err = "level '" + lvl + "' has no writer"; return OSM_B200_ERR_INVALID; }
    std::vector<const osm_b200_component *> stageComps;
    // a temporal stage reading several levels (reader.dmLevel = a;b) sees their implicit concat
    // (core/dataReader.cpp:360-444), i.e. it behaves like stage(cVectorConcat(a, b))
    bool multi = false;
    while (c && (c->type == OSM_B200_C_DELTAREGRESSION || c->type == OSM_B200_C_CONTOURSMOOTHER || c->type == OSM_B200_C_FULLINPUTMEAN)) {
      stageComps.insert(stageComps.begin(), c);
      if (c->n_inputs > 1) { multi = true; break; }
      c = single_input(c);
    }
    if (!c) { err = "broken temporal chain below level '" + lvl + "'"; return OSM_B200_ERR_INVALID; }
    stageComps.insert(stageComps.end(), above.begin(), above.end());
    if (c->type == OSM_B200_C_DATASELECTOR && !multi) {
      // cDataSelector (core/dataSelector.cpp:296-366): picks elements of its (implicitly concatenated) input levels by
      // exact name, in the order of `selected`, each as a single-element field.  Element-wise like the temporal
      // stages, so stage(select(x)) = select(stage(x)): the leaves below carry the stages, the selection is applied
      // to the output groups afterwards (selector scopes).
      const auto &q = c->u.dataselector;
      if (!q.elementMode) { err = "cDataSelector.elementMode=0 is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      if (q.nSelected < 1) { err = "cDataSelector: no elements selected"; return OSM_B200_ERR_INVALID; }
      if (arraysOnly) { err = "cDataSelector below a cVectorConcat that drops single-element fields is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      if (inSelector) { err = "nested cDataSelector levels are not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      if (c->n_inputs < 1) { err = "cDataSelector without inputs"; return OSM_B200_ERR_INVALID; }
      const size_t l0 = leaves.size();
      inSelector = true;
      for (int i = 0; i < c->n_inputs; i++) {
        osm_b200_status s2 = expand(c->reader_dmLevel[i], stageComps, depth + 1, false);
        if (s2 != OSM_B200_OK) { inSelector = false; return s2; }
      }
      inSelector = false;
      if (!stageComps.empty() && c->n_inputs > 1) concatChecks.push_back({l0, leaves.size(), stageComps.size()});
      selScopes.push_back(SelScope{c, l0, leaves.size(), stageComps.size()});
      return OSM_B200_OK;
    }
    if (c->type == OSM_B200_C_VALBASEDSELECTOR && !multi &&
        !(c->n_inputs == 2 && R.prod(c->reader_dmLevel[1]) && R.prod(c->reader_dmLevel[1])->type == OSM_B200_C_PITCHSMOOTHERVITERBI)) {
      // cValbasedSelector over [selector level ; data levels ...] with zeroVec = 1 (GeMAPSv01b_core.lld.conf.inc:395-433: formants /
      // spectral parameters of the voiced resp. unvoiced frames): every frame is written, its elements either copied or set to
      // outputVal, so the gate works element by element like the temporal stages above it.  The pitch chain's own selector
      // ([energy ; cPitchSmootherViterbi level]) is part of the SHS pitch op (get_op).
      const auto &q = c->u.valbasedselector;
      if (c->n_inputs < 2) { err = "cValbasedSelector must read at least two levels: selector;data"; return OSM_B200_ERR_UNSUPPORTED; }
      if (q.idx != 0 || !q.removeIdx || !q.zeroVec || q.adaptiveThreshold) { err = "cValbasedSelector: only idx=0, removeIdx=1, zeroVec=1, adaptiveThreshold=0 are supported"; return OSM_B200_ERR_UNSUPPORTED; }
      if (arraysOnly) { err = "cValbasedSelector below a cVectorConcat that drops single-element fields is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      for (const GateScope &gs : gateScopes) if (gs.l1 == (size_t)-1) { err = "nested cValbasedSelector levels are not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      const size_t l0 = leaves.size();
      gateScopes.push_back(GateScope{c, l0, (size_t)-1, stageComps.size()});
      const size_t me = gateScopes.size() - 1;
      for (int i = 1; i < c->n_inputs; i++) {
        osm_b200_status s2 = expand(c->reader_dmLevel[i], stageComps, depth + 1, false);
        if (s2 != OSM_B200_OK) return s2;
      }
      gateScopes[me].l1 = leaves.size();
      for (size_t l = l0; l < leaves.size(); l++)
        if (leaves[l].stages.size() != stageComps.size()) { err = "cValbasedSelector reading an already smoothed level is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      if (!stageComps.empty() && c->n_inputs > 2) concatChecks.push_back({l0, leaves.size(), stageComps.size()});
      return OSM_B200_OK;
    }
    if (c->type == OSM_B200_C_VECTORCONCAT || multi) {
      if (c->n_inputs < 1) { err = "cVectorConcat without inputs"; return OSM_B200_ERR_INVALID; }
      // a real cVectorConcat is a cVectorProcessor: with processArrayFields=1 it drops single-element
      // fields unless includeSingleElementFields=1 (core/vectorProcessor.cpp:196-243)
      if (!multi && c->u.vectorconcat.processArrayFields == 1 && !c->u.vectorconcat.includeSingleElementFields) arraysOnly = true;
      if (!multi && c->u.vectorconcat.processArrayFields == 2) { err = "cVectorConcat.processArrayFields=2 is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
      const size_t g0 = leaves.size();
      for (int i = 0; i < c->n_inputs; i++) {
        osm_b200_status s2 = expand(c->reader_dmLevel[i], stageComps, depth + 1, arraysOnly);
        if (s2 != OSM_B200_OK) return s2;
      }
      if (!stageComps.empty()) concatChecks.push_back({g0, leaves.size(), stageComps.size()});
      return OSM_B200_OK;
    }
    leaves.push_back(Leaf{c, stageComps, arraysOnly});
    return OSM_B200_OK;
  };
  {
    osm_b200_status s0 = expand(outputLevel, {}, 0, false);
    if (s0 != OSM_B200_OK) return s0;
  }

  // static feature producer of component c (created on first use)
  std::function<osm_b200_status(const osm_b200_component *, int &)> get_op =
    [&](const osm_b200_component *c, int &opIdx) -> osm_b200_status {
    {
      auto it = staticOpOf.find(c);
      if (it != staticOpOf.end()) { opIdx = it->second; return OSM_B200_OK; }
      StaticOp op;
      ChainInfo ci;
      std::string base;
      auto wave_name = [&]() { return std::string(ci.wav->u.wavesource.outFieldName[0] ? ci.wav->u.wavesource.outFieldName : "pcm"); };
      if (c->type == OSM_B200_C_MELSPEC) {
        // the band level itself as an output (log-mel spectrogram style graphs): the band values behind the kernel's filterbank
        // phase, i.e. the cPlp back end with every stage switched off (doLog = doAud = doInvLog = doIDFT = doLP = doLpToCeps = 0
        // hands the bands through, lldcore/plp.cpp:416-593) -- the band sums and the htk scaling are cMelspec's own (a-7)
        const osm_b200_component *mel = c;
        if (!resolve_mag_chain(single_input(mel), ci)) return OSM_B200_ERR_UNSUPPORTED;
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        const FrontEnd &fe = d.streams[op.stream].fe;
        base = name_append_auto(*ci.mag, wave_name(), "fftMag");        // dspcore/fftmagphase.cpp:154
        const auto &melp = mel->u.melspec;
        if (melp.nBands < 1 || melp.nBands > 64 || melp.nBands >= fe.nBins) { err = "cMelspec.nBands out of range"; return OSM_B200_ERR_UNSUPPORTED; }
        MelBank mb;
        build_mel(melp, fe.nBins, fe.fftFrameSizeSec, mb);
        d.mels.push_back(mb);
        osm_b200_plp off;
        memset(&off, 0, sizeof off);
        off.firstCC = 0; off.lastCC = -1; off.nCeps = -1; off.compression = 1.0;
        op.kind = SOP_PLP;
        if (!build_plp(off, d.mels.back(), fe.frameStepSec, op.plp, err)) return OSM_B200_ERR_UNSUPPORTED;
        op.plp.melIdx = (int)d.mels.size() - 1;
        op.nOut = op.plp.nOut;
        if (op.nOut != melp.nBands) { err = "internal: cMelspec pass-through width"; return OSM_B200_ERR_INVALID; }
        FieldName fn;
        fn.name = name_append_auto(*mel, base, nullptr);                // lldcore/melspec.cpp:175-178 (no default nameAppend)
        fn.n = op.nOut; fn.arrNameOffset = 0;
        op.fields.push_back(fn);
      } else if (c->type == OSM_B200_C_MFCC || c->type == OSM_B200_C_PLP) {
        const osm_b200_component *mel = single_input(c);
        if (!mel || mel->type != OSM_B200_C_MELSPEC) { err = "cMfcc / cPlp must read a cMelspec level"; return OSM_B200_ERR_UNSUPPORTED; }
        if (!resolve_mag_chain(single_input(mel), ci)) return OSM_B200_ERR_UNSUPPORTED;
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        const FrontEnd &fe = d.streams[op.stream].fe;
        // field base name: outFieldName -> (pe/win/fft keep) -> fftMag -> melspec keeps
        base = name_append_auto(*ci.mag, wave_name(), "fftMag");        // dspcore/fftmagphase.cpp:154
        base = name_append_auto(*mel, base, nullptr);
        const auto &melp = mel->u.melspec;
        if (melp.nBands < 1 || melp.nBands > 64 || melp.nBands >= fe.nBins) { err = "cMelspec.nBands out of range"; return OSM_B200_ERR_UNSUPPORTED; }
        MelBank mb;
        build_mel(melp, fe.nBins, fe.fftFrameSizeSec, mb);
        d.mels.push_back(mb);
        FieldName fn;
        if (c->type == OSM_B200_C_MFCC) {
          op.kind = SOP_MFCC;
          const auto &mfp = c->u.mfcc;
          if (mfp.lastMfcc < mfp.firstMfcc || mfp.firstMfcc < 0 || mfp.lastMfcc >= melp.nBands) { err = "cMfcc: bad firstMfcc/lastMfcc"; return OSM_B200_ERR_INVALID; }
          build_mfcc(mfp, melp.nBands, op.mfcc);
          op.mfcc.melIdx = (int)d.mels.size() - 1;
          op.nOut = op.mfcc.nMfcc;
          fn.name = name_append_auto(*c, base, nullptr);                // lldcore/mfcc.cpp:120-128
          fn.n = op.nOut; fn.arrNameOffset = op.mfcc.first;             // lldcore/mfcc.cpp:125
        } else {
          op.kind = SOP_PLP;
          if (!build_plp(c->u.plp, d.mels.back(), fe.frameStepSec, op.plp, err)) return OSM_B200_ERR_UNSUPPORTED;
          op.plp.melIdx = (int)d.mels.size() - 1;
          op.nOut = op.plp.nOut;
          // lldcore/plp.cpp:232-267 replaces the field name, then cVectorProcessor appends nameAppend
          const int ra = op.plp.rasta;
          const char *fixed = op.plp.doLpToCeps ? (ra ? "RASTAPlpCC" : "PlpCC")
                            : (op.plp.doLP ? (ra == 1 ? "RASTAPlpc" : (ra == 2 ? "newRASTAPlpc" : "Plpc"))
                            : (op.plp.doIDFT ? "audAutoCor" : "audSpec"));
          fn.name = name_append_auto(*c, fixed, nullptr);
          fn.n = op.nOut; fn.arrNameOffset = 0;
        }
        op.fields.push_back(fn);
      } else if (c->type == OSM_B200_C_SPECTRAL) {
        if (!resolve_mag_chain(single_input(c), ci)) return OSM_B200_ERR_UNSUPPORTED;
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        const FrontEnd &fe = d.streams[op.stream].fe;
        op.kind = SOP_SPECTRAL;
        if (!build_spectral(c->u.spectral, fe.nBins, fe.fftFrameSizeSec, op.spectral, err)) return OSM_B200_ERR_UNSUPPORTED;
        op.nOut = op.spectral.nOut;
        base = name_append_auto(*ci.mag, wave_name(), "fftMag");
        // element names, lldcore/spectral.cpp:378-584
        const auto &sc = c->u.spectral;
        const bool lg = sc.useLogSpectrum != 0;
        auto add = [&](const std::string &suffix) { FieldName f; f.name = base + "_" + suffix; op.fields.push_back(f); };
        for (int i = 0; i < sc.nBands; i++)
          if ((long)sc.bandLo[i] >= 0 && (long)sc.bandHi[i] > 0) { snprintf(buf, sizeof buf, "%s%ld-%ld", lg ? "logFband" : "fband", (long)sc.bandLo[i], (long)sc.bandHi[i]); add(buf); }
        for (int i = 0; i < sc.nSlopes; i++)
          if ((long)sc.slopeLo[i] >= 0 && (long)sc.slopeHi[i] > 0) { snprintf(buf, sizeof buf, "%s%ld-%ld", lg ? "logSpectralSlopeOfBand" : "spectralSlopeOfBand", (long)sc.slopeLo[i], (long)sc.slopeHi[i]); add(buf); }
        if (sc.alphaRatio) add(lg ? "alphaRatioDB" : "alphaRatio");
        if (sc.hammarbergIndex) add(lg ? "hammarbergIndexDB" : "hammarbergIndex");
        for (size_t i = 0; i < op.spectral.rollOff.size(); i++) { snprintf(buf, sizeof buf, "spectralRollOff%.1f", op.spectral.rollOff[i] * 100.0); add(buf); }
        if (sc.flux) add("spectralFlux");
        if (sc.centroid) add(lg ? "logSpectralCentroid" : "spectralCentroid");
        if (sc.maxPos) add("spectralMaxPos");
        if (sc.minPos) add("spectralMinPos");
        if (sc.entropy) add(lg ? "logSpectralEntropy" : "spectralEntropy");
        if (sc.standardDeviation) add(lg ? "logSpectralStdDev" : "spectralStdDev");
        if (sc.variance) add(lg ? "logSpectralVariance" : "spectralVariance");
        if (sc.skewness) add(lg ? "logSpectralSkewness" : "spectralSkewness");
        if (sc.kurtosis) add(lg ? "logSpectralKurtosis" : "spectralKurtosis");
        if (sc.slope) add(lg ? "logSpectralSlope" : "spectralSlope");
        if (sc.sharpness) add("psySharpness");
        if (sc.harmonicity) add(lg ? "logSpectralHarmonicity" : "spectralHarmonicity");
        if (sc.flatness) add(lg ? "logSpectralFlatness" : "spectralFlatness");
        if ((int)op.fields.size() != op.nOut) { err = "internal: cSpectral name/element mismatch"; return OSM_B200_ERR_INVALID; }
      } else if (c->type == OSM_B200_C_PITCHACF) {
        // reader.dmLevel = <acf level>;<cepstrum level> (lldcore/pitchACF.cpp:148-152), both cAcf
        // instances on the same magnitude level
        if (c->n_inputs != 2) { err = "cPitchACF must read two levels: acf;cepstrum"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *a = R.prod(c->reader_dmLevel[0]), *b = R.prod(c->reader_dmLevel[1]);
        if (!a || !b || a->type != OSM_B200_C_ACF || b->type != OSM_B200_C_ACF) { err = "cPitchACF inputs must be cAcf levels"; return OSM_B200_ERR_UNSUPPORTED; }
        if (a->u.acf.cepstrum || !b->u.acf.cepstrum) { err = "cPitchACF expects [acf ; cepstrum] in this order"; return OSM_B200_ERR_UNSUPPORTED; }
        for (const osm_b200_component *x : {a, b}) {
          const auto &q = x->u.acf;
          if (q.inverse || q.cosLifterCepstrum || q.oldCompatCepstrum || !q.symmetricData) { err = "cAcf: inverse / cosLifterCepstrum / oldCompatCepstrum / symmetricData=0 are not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        }
        if (single_input(a) != single_input(b)) { err = "both cAcf instances must read the same cFFTmagphase level"; return OSM_B200_ERR_UNSUPPORTED; }
        if (!resolve_mag_chain(single_input(a), ci)) return OSM_B200_ERR_UNSUPPORTED;
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        const FrontEnd &fe = d.streams[op.stream].fe;
        op.kind = SOP_PITCHACF;
        PitchAcfOp &po = op.pitch;
        po.acfUsePower = a->u.acf.usePower != 0; po.cepUsePower = b->u.acf.usePower != 0;
        po.absCepstrum = b->u.acf.absCepstrum != 0;
        po.normOutput = a->u.acf.acfCepsNormOutput != 0;
        if ((b->u.acf.acfCepsNormOutput != 0) != po.normOutput) { err = "cAcf.acfCepsNormOutput must agree on both instances"; return OSM_B200_ERR_UNSUPPORTED; }
        const auto &pp = c->u.pitchacf;
        po.maxPitch = pp.maxPitch < 0.0 ? 0.0 : pp.maxPitch;                       // lldcore/pitchACF.cpp:101-102
        po.voicingCutoff = pp.voicingCutoff > 1.0 ? 1.0 : (pp.voicingCutoff < 0.0 ? 0.0 : pp.voicingCutoff);
        po.fsSec = (float)fe.fftFrameSizeSec;                                       // :107-110
        po.voiceProb = pp.voiceProb != 0; po.voiceQual = pp.voiceQual != 0; po.HNR = pp.HNR != 0; po.HNRdB = pp.HNRdB != 0;
        po.linHNR = pp.linHNR != 0; po.F0 = pp.F0 != 0; po.F0raw = pp.F0raw != 0; po.F0env = pp.F0env != 0;
        auto add = [&](const char *nm) { FieldName f; f.name = nm; op.fields.push_back(f); };   // :112-121
        if (po.voiceProb) add("voiceProb");
        if (po.HNR) add("HNR");
        if (po.HNRdB) add("HNRdBacf");
        if (po.linHNR) add("linearHNRacf");
        if (po.voiceQual) add("voiceQual");
        if (po.F0) add("F0");
        if (po.F0raw) add("F0raw");
        if (po.F0env) add("F0env");
        po.nOut = (int)op.fields.size();
        op.nOut = po.nOut;
        if (op.nOut < 1) { err = "cPitchACF produces no output"; return OSM_B200_ERR_INVALID; }
      } else if (c->type == OSM_B200_C_ENERGY || c->type == OSM_B200_C_MZCR || c->type == OSM_B200_C_INTENSITY) {
        const osm_b200_component *in = single_input(c);
        if (!resolve_time_chain(in, ci)) return OSM_B200_ERR_UNSUPPORTED;
        op.windowed = ci.win != nullptr;
        if (ci.pe && !ci.win) { err = "cEnergy / cMZcr reading a pre-emphasised, un-windowed level is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        osm_b200_status s2 = get_stream(ci, false, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        base = wave_name();
        if (c->type == OSM_B200_C_ENERGY) {
          op.kind = SOP_ENERGY;
          build_energy(c->u.energy, op.energy);
          op.nOut = op.energy.nOut;
          // lldcore/energy.cpp:97-125: addNameAppendFieldAuto(name, "RMS"|"SQUARED"|"LOG")
          if (op.energy.rms) { FieldName f; f.name = name_append_auto(*c, base, "RMS"); op.fields.push_back(f); }
          if (op.energy.energy2) { FieldName f; f.name = name_append_auto(*c, base, "SQUARED"); op.fields.push_back(f); }
          if (op.energy.lg) { FieldName f; f.name = name_append_auto(*c, base, "LOG"); op.fields.push_back(f); }
        } else if (c->type == OSM_B200_C_INTENSITY) {
          op.kind = SOP_INTENSITY;
          IntensityOp &io = op.intensity;
          io.intensity = c->u.intensity.intensity != 0; io.loudness = c->u.intensity.loudness != 0;
          io.nOut = (io.intensity ? 1 : 0) + (io.loudness ? 1 : 0);
          op.nOut = io.nOut;
          // Hamming window of the frame length (smileutil/smileUtil.c:1291-1303), summed in index order
          const int N = d.streams[op.stream].fe.frameSize;
          io.winSum = 0.0;
          for (int j = 0; j < N; j++) {
            const double w = 0.54 - 0.46 * cos((2.0 * M_PI * (double)j) / ((double)N - 1.0));
            if (j < 2) io.w[j] = w;
            io.winSum += w;
          }
          if (io.winSum <= 0.0) io.winSum = 1.0;
          if (io.intensity) { FieldName f; f.name = name_append_auto(*c, base, "intensity"); op.fields.push_back(f); }   // lldcore/intensity.cpp:91-92
          if (io.loudness) { FieldName f; f.name = name_append_auto(*c, base, "loudness"); op.fields.push_back(f); }
        } else {
          op.kind = SOP_MZCR;
          build_mzcr(c->u.mzcr, op.mzcr);
          op.nOut = op.mzcr.nOut;
          auto add = [&](const char *suffix) { FieldName f; f.name = base + "_" + suffix; op.fields.push_back(f); };   // lldcore/mzcr.cpp:68-100
          if (op.mzcr.zcr) add("zcr");
          if (op.mzcr.mcr) add("mcr");
          if (op.mzcr.amax) add("absmax");
          if (op.mzcr.maxmin) { add("max"); add("min"); }
          if (op.mzcr.dc) add("dc");
        }
        if (op.nOut < 1) { err = "component produces no output"; return OSM_B200_ERR_INVALID; }
      } else if (c->type == OSM_B200_C_FFTMAGPHASE) {
        // the magnitude level itself as an output (config/spectrum/spectrogram.conf): nBins elements
        if (!resolve_mag_chain(c, ci, true)) return OSM_B200_ERR_UNSUPPORTED;
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        op.kind = SOP_MAG;
        op.nOut = d.streams[op.stream].fe.nBins;
        const auto &mp = c->u.fftmagphase;
        const bool norm = mp.normalise || mp.dBpsd;                          // dBpsd implies normalise (:92)
        op.magMode = mp.dBpsd ? 4 : (norm && mp.power ? 3 : (mp.power ? 2 : (norm ? 1 : 0)));
        op.magDbNorm = (float)mp.dBpnorm;
        op.magMinDb = (float)mp.mindBp;
        if (op.magMinDb - op.magDbNorm < -120.0f) op.magMinDb = -120.0f + op.magDbNorm;      // :95-98
        const char *fixed = mp.dBpsd ? "fftMag_dBsplPSD" : (mp.power && !norm ? "fftMag_PowSpec" : (mp.power ? "fftMag_PowSpecDens" : (norm ? "fftMag_SpecDens" : "fftMag")));   // :141-155
        FieldName fn;
        fn.name = name_append_auto(*c, wave_name(), fixed);
        fn.n = op.nOut; fn.arrNameOffset = 0;
        op.fields.push_back(fn);
      } else if (c->type == OSM_B200_C_VECTOROPERATION) {
        // n -> 1 reduction of another static level (other/vectorOperation.cpp:475-481, names :226-249)
        if (c->u.vectoroperation.operation != 0) { err = "cVectorOperation: only operation=ll1 is supported"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *in = single_input(c);
        if (!in) { err = "cVectorOperation must read exactly one level"; return OSM_B200_ERR_UNSUPPORTED; }
        int src = -1;
        osm_b200_status s2 = get_op(in, src);
        if (s2 != OSM_B200_OK) return s2;
        if (d.ops[src].fields.size() != 1) { err = "cVectorOperation: the input level must hold exactly one field"; return OSM_B200_ERR_UNSUPPORTED; }
        op.kind = SOP_VECOP;
        op.srcOp = src;
        op.stream = d.ops[src].stream;
        op.nOut = 1;
        FieldName fn;
        osm_b200_component named = *c;
        if (!named.nameAppend[0]) snprintf(named.nameAppend, sizeof named.nameAppend, "%s", "lengthL1norm");
        const std::string inName = c->u.vectoroperation.nameBase[0] ? std::string(c->u.vectoroperation.nameBase) : d.ops[src].fields[0].name;
        fn.name = name_append_auto(named, inName, nullptr);
        op.fields.push_back(fn);
      } else if (c->type == OSM_B200_C_FORMANTLPC) {
        // cFormantLpc <- cLpc <- cSpecResample <- cTransformFFT <- cWindower chain (GeMAPSv01b_core.lld.conf.inc:250-286)
        const osm_b200_component *lpc = single_input(c);
        if (!lpc || lpc->type != OSM_B200_C_LPC) { err = "cFormantLpc must read a cLpc level"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *rsm = single_input(lpc);
        if (!rsm || rsm->type != OSM_B200_C_SPECRESAMPLE) { err = "cLpc must read a cSpecResample level"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *fft = single_input(rsm);
        if (!fft || fft->type != OSM_B200_C_TRANSFORMFFT || fft->u.transformfft.inverse) { err = "cSpecResample must read a (forward) cTransformFFT level"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *w = single_input(fft);
        if (!w || w->type != OSM_B200_C_WINDOWER) { err = "cTransformFFT must read a cWindower level"; return OSM_B200_ERR_UNSUPPORTED; }
        if (!resolve_time_chain(w, ci)) return OSM_B200_ERR_UNSUPPORTED;
        op.windowed = true;
        osm_b200_status s2 = get_stream(ci, false, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        op.kind = SOP_FORMANT;
        if (!build_formant(rsm->u.specresample, lpc->u.lpc, c->u.formantlpc, d.streams[op.stream].fe,
                           fft->u.transformfft.zeroPadSymmetric != 0, op.formant, err)) return OSM_B200_ERR_UNSUPPORTED;
        op.nOut = op.formant.nOut;
        // lld/formantLpc.cpp:113-136: fixed field names, array indices start at 1
        if (op.formant.saveNValid) { FieldName f; f.name = "nFormants"; op.fields.push_back(f); }
        if (op.formant.saveFormants) { FieldName f; f.name = "formantFreqLpc"; f.n = op.formant.nFormants; f.arrNameOffset = 1; op.fields.push_back(f); }
        if (op.formant.saveBandwidths) { FieldName f; f.name = "formantBandwidthLpc"; f.n = op.formant.nFormants; f.arrNameOffset = 1; op.fields.push_back(f); }
      } else if (c->type == OSM_B200_C_HARMONICS) {
        // cHarmonics reads [pitch level ; formant level ; magnitude level] through one multi-level reader
        // (GeMAPSv01b_core.lld.conf.inc:289-318) and looks its inputs up by name (lld/harmonics.cpp:258-300)
        const auto &q = c->u.harmonics;
        const osm_b200_component *pit = nullptr, *fmt = nullptr, *mg = nullptr;
        for (int i = 0; i < c->n_inputs; i++) {
          const osm_b200_component *x = R.prod(c->reader_dmLevel[i]);
          if (!x) { err = std::string("level '") + c->reader_dmLevel[i] + "' has no writer"; return OSM_B200_ERR_INVALID; }
          if (x->type == OSM_B200_C_VALBASEDSELECTOR || x->type == OSM_B200_C_PITCHSMOOTHERVITERBI) pit = x;
          else if (x->type == OSM_B200_C_FORMANTLPC) fmt = x;
          else if (x->type == OSM_B200_C_FFTMAGPHASE) mg = x;
          else { err = "cHarmonics: inputs must be a Viterbi-smoothed pitch level, a cFormantLpc level and a cFFTmagphase level"; return OSM_B200_ERR_UNSUPPORTED; }
        }
        if (!pit || !mg || c->n_inputs > 3) { err = "cHarmonics needs a pitch level and a magnitude level (and optionally a formant level)"; return OSM_B200_ERR_UNSUPPORTED; }
        if (q.nHarmonicMagnitudes > 0 || q.outputLinearMagnitudes || q.harmonicDifferencesRatioLinear || q.formantAmplitudesLinear || q.computeAcfHnrLinear) {
          err = "cHarmonics: harmonic magnitudes / linear outputs are not supported (log differences, log formant amplitudes, HNR in dB only)"; return OSM_B200_ERR_UNSUPPORTED;
        }
        if (!resolve_mag_chain(mg, ci)) return OSM_B200_ERR_UNSUPPORTED;
        HarmonicsOp &ho = op.harmonics;
        osm_b200_status s3 = get_op(pit, ho.pitchOp);
        if (s3 != OSM_B200_OK) return s3;
        if (fmt) { s3 = get_op(fmt, ho.formantOp); if (s3 != OSM_B200_OK) return s3; }
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        const FrontEnd &fe = d.streams[op.stream].fe;
        {
          const FrontEnd &fp = d.streams[d.ops[ho.pitchOp].stream].fe;
          if (fp.frameSize != fe.frameSize || fp.frameStep != fe.frameStep) { err = "cHarmonics: the pitch level and the magnitude level must share the frame geometry"; return OSM_B200_ERR_UNSUPPORTED; }
          if (fmt) {
            const FrontEnd &ff = d.streams[d.ops[ho.formantOp].stream].fe;
            if (ff.frameStep != fe.frameStep || ff.frameSize > fe.frameSize) { err = "cHarmonics: the formant level must have the same frame step and frames no longer than the magnitude level's"; return OSM_B200_ERR_UNSUPPORTED; }
          }
        }
        // the magnitude field (name of the cFFTmagphase level, dspcore/fftmagphase.cpp:154)
        {
          const std::string magName = name_append_auto(*mg, wave_name(), "fftMag");
          const bool ok = q.magSpecFieldNameIsFull ? magName == q.magSpecFieldName : magName.find(q.magSpecFieldName) != std::string::npos;
          if (!ok) { err = "cHarmonics: magSpecFieldName '" + std::string(q.magSpecFieldName) + "' does not match the magnitude level's field '" + magName + "'"; return OSM_B200_ERR_INVALID; }
        }
        auto find_field = [&](int opI, const char *name, bool full, int &colOut, int &nOutF) -> bool {
          int col = d.ops[opI].outCol;
          for (const FieldName &f : d.ops[opI].fields) {
            if (full ? f.name == name : f.name.find(name) != std::string::npos) { colOut = col; nOutF = f.n; return true; }
            col += f.n;
          }
          return false;
        };
        int nF0 = 0;
        if (!find_field(ho.pitchOp, q.f0ElementName, q.f0ElementNameIsFull != 0, ho.f0Col, nF0) || nF0 != 1) {
          err = "cHarmonics: f0ElementName '" + std::string(q.f0ElementName) + "' not found in the pitch level"; return OSM_B200_ERR_INVALID;
        }
        bool fa = q.formantAmplitudes != 0 && q.formantAmplitudesLogRel != 0;
        bool haveFormantDiff = false;
        if (fmt && q.formantFrequencyFieldName[0]) {
          if (!find_field(ho.formantOp, q.formantFrequencyFieldName, q.formantFrequencyFieldNameIsFull != 0, ho.fmtCol, ho.nFmt)) {
            err = "cHarmonics: formantFrequencyFieldName '" + std::string(q.formantFrequencyFieldName) + "' not found in the formant level"; return OSM_B200_ERR_INVALID;
          }
          int bc = 0, bn = 0;
          if (!q.formantBandwidthFieldName[0] || !find_field(ho.formantOp, q.formantBandwidthFieldName, q.formantBandwidthFieldNameIsFull != 0, bc, bn) || bn != ho.nFmt) {
            err = "cHarmonics: formantBandwidthFieldName must name the bandwidth field of the formant level (lld/harmonics.cpp:270-296)"; return OSM_B200_ERR_UNSUPPORTED;
          }
          if (ho.nFmt > 8) { err = "cHarmonics: more than 8 formants"; return OSM_B200_ERR_UNSUPPORTED; }
        } else fa = false;
        // harmonicDifferences: "H<i>-H<j>", "H<i>-A<k>", ... (:84-160); A0 = the fundamental
        int maxHarm = 0;
        for (int i = 0; i < q.nHarmonicDifferences && q.harmonicDifferencesLog; i++) {
          const char *t = q.harmonicDifferences[i];
          const char *dash = strchr(t, '-');
          if (!dash || dash == t) { err = std::string("cHarmonics: cannot parse harmonic difference '") + t + "'"; return OSM_B200_ERR_INVALID; }
          int part[2][2];                                            // {formant, idx}
          const char *ps[2] = {t, dash + 1};
          for (int k = 0; k < 2; k++) {
            char *ep = nullptr;
            const long r = strtol(ps[k] + 1, &ep, 10);
            if (ep == ps[k] + 1 || (ps[k][0] != 'H' && ps[k][0] != 'A')) { err = std::string("cHarmonics: cannot parse harmonic difference '") + t + "'"; return OSM_B200_ERR_INVALID; }
            if (ps[k][0] == 'H') { part[k][0] = -1; part[k][1] = (int)r; if (r > maxHarm) maxHarm = (int)r; }
            else if (r == 0) { part[k][0] = -1; part[k][1] = 0; }
            else { part[k][0] = (int)r; part[k][1] = -1; haveFormantDiff = true; }
          }
          ho.diffs.insert(ho.diffs.end(), {part[0][0], part[0][1], part[1][0], part[1][1]});
        }
        if (haveFormantDiff && ho.nFmt == 0) ho.diffs.clear();       // :296-300: disabled without a formant level
        ho.nHarm = q.nHarmonics;
        if (ho.nHarm < q.nHarmonicMagnitudes + q.firstHarmonicMagnitude + 1) ho.nHarm = q.nHarmonicMagnitudes + q.firstHarmonicMagnitude + 1;   // :212-217
        if (ho.nHarm < maxHarm + 1) ho.nHarm = maxHarm + 1;
        if (ho.nHarm < 2 || ho.nHarm > 128) { err = "cHarmonics.nHarmonics must be in 2..128"; return OSM_B200_ERR_UNSUPPORTED; }
        ho.hnr = q.computeAcfHnrLogdB != 0;
        ho.fa = fa;
        if (fa) {                                                    // :341-352
          ho.faStart = q.formantAmplitudesStart < 0 ? 0 : q.formantAmplitudesStart;
          ho.faEnd = q.formantAmplitudesEnd == -1 ? ho.nFmt : std::min(q.formantAmplitudesEnd, ho.nFmt);
          if (ho.faEnd < ho.faStart) ho.fa = false;
          else if (ho.faStart < 1) { err = "cHarmonics.formantAmplitudesStart=0 is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        }
        ho.floorUnvoiced = (float)q.logRelValueFloorUnvoiced;
        ho.nb = fe.nBins;
        ho.binHz = 1.0 / fe.fftFrameSizeSec;                         // dspcore/transformFft.cpp:111-115
        op.kind = SOP_HARMONICS;
        if (ho.hnr) { FieldName f; f.name = "HarmonicsToNoiseRatioACFLogdB"; op.fields.push_back(f); }   // :236-240
        for (size_t i = 0; i < ho.diffs.size() / 4; i++) { FieldName f; f.name = std::string("HarmonicDifferenceLogRel") + q.harmonicDifferences[i]; op.fields.push_back(f); }
        if (ho.fa) { FieldName f; f.name = "FormantAmplitudeByMaxHarmonicLogRelF0"; f.n = ho.faEnd - ho.faStart + 1; f.arrNameOffset = ho.faStart; op.fields.push_back(f); }
        ho.nOut = 0;
        for (const FieldName &f : op.fields) ho.nOut += f.n;
        op.nOut = ho.nOut;
        if (op.nOut < 1) { err = "cHarmonics produces no output"; return OSM_B200_ERR_INVALID; }
      } else if (c->type == OSM_B200_C_VALBASEDSELECTOR || c->type == OSM_B200_C_PITCHSMOOTHERVITERBI) {
        // [cValbasedSelector <-] cPitchSmootherViterbi <- cPitchShs <- cSpecScale <- cFFTmagphase chain
        const osm_b200_component *vit = c, *selSrc = nullptr;
        if (c->type == OSM_B200_C_VALBASEDSELECTOR) {
          const auto &q = c->u.valbasedselector;
          if (c->n_inputs != 2) { err = "cValbasedSelector must read two levels: selector;data"; return OSM_B200_ERR_UNSUPPORTED; }
          if (q.idx != 0 || !q.removeIdx || !q.zeroVec || q.adaptiveThreshold) { err = "cValbasedSelector: only idx=0, removeIdx=1, zeroVec=1, adaptiveThreshold=0 are supported"; return OSM_B200_ERR_UNSUPPORTED; }
          selSrc = R.prod(c->reader_dmLevel[0]);
          vit = R.prod(c->reader_dmLevel[1]);
          if (!selSrc || !vit || vit->type != OSM_B200_C_PITCHSMOOTHERVITERBI) { err = "cValbasedSelector: the data level must come from cPitchSmootherViterbi"; return OSM_B200_ERR_UNSUPPORTED; }
        }
        const osm_b200_component *shs = single_input(vit);
        if (!shs || shs->type != OSM_B200_C_PITCHSHS) { err = "cPitchSmootherViterbi must read a cPitchShs level"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *scl = single_input(shs);
        if (!scl || scl->type != OSM_B200_C_SPECSCALE) { err = "cPitchShs must read a cSpecScale level"; return OSM_B200_ERR_UNSUPPORTED; }
        if (!resolve_mag_chain(single_input(scl), ci)) return OSM_B200_ERR_UNSUPPORTED;
        int selOp = -1;
        if (selSrc) {
          osm_b200_status s3 = get_op(selSrc, selOp);                   // e.g. cEnergy (rms) on the same frames
          if (s3 != OSM_B200_OK) return s3;
          if (d.ops[selOp].nOut != 1) { err = "cValbasedSelector: the selector level must hold exactly one element"; return OSM_B200_ERR_UNSUPPORTED; }
        }
        osm_b200_status s2 = get_stream(ci, true, op.stream);
        if (s2 != OSM_B200_OK) return s2;
        const FrontEnd &fe = d.streams[op.stream].fe;
        if (selOp >= 0) {
          const FrontEnd &fs = d.streams[d.ops[selOp].stream].fe;
          if (fs.frameSize != fe.frameSize || fs.frameStep != fe.frameStep) { err = "cValbasedSelector: selector and data levels must share the frame geometry"; return OSM_B200_ERR_UNSUPPORTED; }
        }
        op.kind = SOP_PITCH;
        if (!build_pitch_chain(scl->u.specscale, shs->u.pitchshs, vit->u.pitchsmootherviterbi, fe.nBins, fe.fftFrameSizeSec, op.chain, err))
          return OSM_B200_ERR_UNSUPPORTED;
        PitchChainOp &pc = op.chain;
        if (selOp >= 0) {
          const auto &q = c->u.valbasedselector;
          pc.hasSel = true; pc.selOp = selOp; pc.selThreshold = (float)q.threshold; pc.selOutputVal = (float)q.outputVal;
          pc.selInvert = q.invert != 0; pc.selAllowEqual = q.allowEqual != 0;
        }
        op.nOut = pc.nOut;
        // field names: lld/pitchSmootherViterbi.cpp:297-316; the selector keeps them (valbasedSelector.cpp:105-118)
        auto add = [&](const char *nm) {
          FieldName f; f.name = nm;
          if (c->type == OSM_B200_C_VALBASEDSELECTOR) f.name = name_append_auto(*c, f.name, nullptr);
          op.fields.push_back(f);
        };
        if (pc.oF0final) add("F0final");
        if (pc.oF0finalLog) add("F0finalLog");
        if (pc.oF0finalEnv) add("F0finEnv");
        if (pc.oF0finalEnvLog) add("F0finEnvLog");
        if (pc.oVClipped) add("voicingFinalClipped");
        if (pc.oVUnclipped) add("voicingFinalUnclipped");
      } else if (c->type == OSM_B200_C_PITCHJITTER) {
        const auto &q = c->u.pitchjitter;
        const osm_b200_component *wv = single_input(c);
        if (!wv || wv->type != OSM_B200_C_WAVESOURCE) { err = "cPitchJitter must read the cWaveSource level"; return OSM_B200_ERR_UNSUPPORTED; }
        const osm_b200_component *f0c = R.prod(q.F0reader_dmLevel);
        if (!f0c) { err = "cPitchJitter: F0reader.dmLevel has no writer"; return OSM_B200_ERR_INVALID; }
        int pOp = -1;
        osm_b200_status s2 = get_op(f0c, pOp);
        if (s2 != OSM_B200_OK) return s2;
        if (d.ops[pOp].kind != SOP_PITCH) { err = "cPitchJitter: the F0 level must come from cPitchSmootherViterbi (optionally through cValbasedSelector)"; return OSM_B200_ERR_UNSUPPORTED; }
        op.kind = SOP_JITTER;
        op.stream = d.ops[pOp].stream;
        JitterOp &jo = op.jitter;
        jo.pitchOp = pOp; jo.f0Col = 0;                                   // lld/pitchJitter.cpp:271-280: unknown field -> element 0
        { int col = 0; for (const auto &f : d.ops[pOp].fields) { if (f.name == q.F0field) { jo.f0Col = col; break; } col += f.n; } }
        jo.searchRangeRel = q.searchRangeRel; jo.lgHNRfloor = q.lgHNRfloor;
        jo.minNumPeriods = q.minNumPeriods < 2 ? 2 : q.minNumPeriods;      // :145-149
        jo.minCC = q.minCC;
        jo.jitterLocal = q.jitterLocal != 0; jo.jitterDDP = q.jitterDDP != 0; jo.jitterLocalEnv = q.jitterLocalEnv != 0; jo.jitterDDPEnv = q.jitterDDPEnv != 0;
        jo.shimmerLocal = q.shimmerLocal != 0; jo.shimmerLocalDB = q.shimmerLocalDB != 0; jo.shimmerLocalEnv = q.shimmerLocalEnv != 0;
        jo.shimmerLocalDBEnv = q.shimmerLocalDBEnv != 0; jo.harmonicERMS = q.harmonicERMS != 0; jo.noiseERMS = q.noiseERMS != 0;
        jo.linearHNR = q.linearHNR != 0; jo.logHNR = q.logHNR != 0; jo.shimmerUseRms = q.shimmerUseRmsAmplitude != 0;
        jo.refinedF0 = q.refinedF0 != 0; jo.srcQualRange = q.sourceQualityRange != 0; jo.srcQualMean = q.sourceQualityMean != 0;
        jo.peakToPeak = q.usePeakToPeakPeriodLength != 0; jo.brokenThresh = q.useBrokenJitterThresh != 0;
        if (q.onlyVoiced) { err = "cPitchJitter.onlyVoiced=1 (variable frame count) is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        if (wv->u.wavesource.nChannels > 1 && !wv->u.wavesource.monoMixdown) { err = "cPitchJitter needs mono input"; return OSM_B200_ERR_UNSUPPORTED; }
        auto add = [&](const std::string &nm) { FieldName f; f.name = nm; op.fields.push_back(f); };   // :283-312
        if (jo.jitterLocal) add("jitterLocal");
        if (jo.jitterDDP) add("jitterDDP");
        if (jo.jitterLocalEnv) add("jitterLocEnv");
        if (jo.jitterDDPEnv) add("jitterDEnv");
        if (jo.shimmerLocal) add("shimmerLocal");
        if (jo.shimmerLocalDB) add("shimmerLocalDB");
        if (jo.shimmerLocalEnv) add("shimmerLocEnv");
        if (jo.shimmerLocalDBEnv) add("shimmerLocDBEnv");
        if (jo.harmonicERMS) add("harmonicERMS");
        if (jo.noiseERMS) add("noiseERMS");
        if (jo.linearHNR) add("linearHNR");
        if (jo.logHNR) add("logHNR");
        if (jo.refinedF0) add(q.F0field[0] ? q.F0field : "F0final");
        if (jo.srcQualMean) add("sourceQualityMean");
        if (jo.srcQualRange) add("sourceQualityRange");
        jo.nOut = (int)op.fields.size();
        op.nOut = jo.nOut;
        if (op.nOut < 1) { err = "cPitchJitter produces no output"; return OSM_B200_ERR_INVALID; }
      } else {
        snprintf(buf, sizeof buf, "component '%s' (%s) is not a supported static LLD producer", c->name, type_name(c->type));
        err = buf; return OSM_B200_ERR_UNSUPPORTED;
      }
      op.outCol = d.nStatic;
      d.nStatic += op.nOut;
      d.ops.push_back(op);
      opIdx = (int)d.ops.size() - 1;
      staticOpOf[c] = opIdx;
    }
    return OSM_B200_OK;
  };

  // ---- cValbasedSelector gates: the selector value must be one column of the SHS pitch level (directly, or through a cDataSelector
  // that picks it: GeMAPSv01b_core.lld.conf.inc:385-391) ----
  struct GateInfo { int col, pitchOp; };
  std::vector<GateInfo> gateInfo;
  for (const GateScope &gs : gateScopes) {
    const osm_b200_component *sl = R.prod(gs.c->reader_dmLevel[0]);
    if (!sl) { err = std::string("level '") + gs.c->reader_dmLevel[0] + "' has no writer"; return OSM_B200_ERR_INVALID; }
    const char *want = nullptr;
    if (sl->type == OSM_B200_C_DATASELECTOR) {
      if (sl->u.dataselector.nSelected != 1 || sl->n_inputs != 1 || !sl->u.dataselector.elementMode) { err = "cValbasedSelector: a cDataSelector selector level must pick exactly one element of one level"; return OSM_B200_ERR_UNSUPPORTED; }
      want = sl->u.dataselector.selected[0];
      sl = R.prod(sl->reader_dmLevel[0]);
      if (!sl) { err = "cValbasedSelector: the selector level has no writer"; return OSM_B200_ERR_INVALID; }
    }
    int so = -1;
    osm_b200_status s3 = get_op(sl, so);
    if (s3 != OSM_B200_OK) return s3;
    if (d.ops[so].kind != SOP_PITCH) { err = "cValbasedSelector: the selector level must come from the SHS pitch chain (cPitchSmootherViterbi)"; return OSM_B200_ERR_UNSUPPORTED; }
    int col = -1, cc = d.ops[so].outCol;
    for (const FieldName &f : d.ops[so].fields) {
      if (f.n == 1 && (want ? f.name == want : d.ops[so].nOut == 1)) { col = cc; break; }
      cc += f.n;
    }
    if (col < 0) { err = want ? std::string("cValbasedSelector: selector element '") + want + "' not found in the pitch level" : std::string("cValbasedSelector: the selector level must hold exactly one element"); return OSM_B200_ERR_UNSUPPORTED; }
    gateInfo.push_back(GateInfo{col, so});
  }

  std::vector<const osm_b200_component *> segComps;     // cDeltaRegression instances with onlyInSegments=1
  for (const Leaf &leaf : leaves) {
    const osm_b200_component *c = leaf.c;
    const std::vector<const osm_b200_component *> &stageComps = leaf.stages;
    int opIdx = -1;
    {
      osm_b200_status so = get_op(c, opIdx);
      if (so != OSM_B200_OK) return so;
    }

    // ---- groups: one per run of consecutive fields that survive the concat's field selection ----
    std::vector<Stage> stages;
    std::vector<FieldName> fields = d.ops[opIdx].fields;
    int segId = -1;
    const size_t leafIdx = (size_t)(&leaf - &leaves[0]);
    long gateIdx = -1;
    for (size_t q = 0; q < gateScopes.size(); q++) if (leafIdx >= gateScopes[q].l0 && leafIdx < gateScopes[q].l1) gateIdx = (long)q;
    if (gateIdx >= 0) for (auto &f : fields) f.name = name_append_auto(*gateScopes[gateIdx].c, f.name, nullptr);   // other/valbasedSelector.cpp:84-97
    long selAbove = -1;                                  // stages above the selector this leaf sits below, -1 = none
    for (const SelScope &sc : selScopes) if (leafIdx >= sc.l0 && leafIdx < sc.l1) selAbove = (long)sc.above;
    auto element_names = [&](const std::vector<FieldName> &fs) {
      std::vector<std::string> out;
      for (const auto &f : fs) {
        if (f.n == 1) out.push_back(f.name);
        else for (int i = 0; i < f.n; i++) { snprintf(buf, sizeof buf, "%s[%d]", f.name.c_str(), i + f.arrNameOffset); out.push_back(buf); }
      }
      return out;
    };
    leafSelNames.push_back({});
    if (selAbove >= 0 && (long)stageComps.size() == selAbove) leafSelNames.back() = element_names(fields);
    size_t stageNo = 0;
    for (const osm_b200_component *s : stageComps) {
      Stage st;
      if (s->type == OSM_B200_C_DELTAREGRESSION) {
        const auto &p = s->u.deltaregression;
        if (p.deltawin < 1 || p.deltawin > 8) { err = "cDeltaRegression.deltawin must be 1..8"; return OSM_B200_ERR_UNSUPPORTED; }
        // flags: bit 0 onlyInSegments, bit 1 relativeDelta, bit 2 absOutput, bit 3 halfWaveRect (dspcore/deltaRegression.cpp:100-168;
        // halfWaveRect wins over absOutput, :157-165)
        const int variant = (p.relativeDelta ? 2 : 0) | (p.halfWaveRect ? 8 : (p.absOutput ? 4 : 0));
        if (variant && p.onlyInSegments) { err = "cDeltaRegression: relativeDelta / absOutput / halfWaveRect together with onlyInSegments are not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        st = Stage{ST_DELTA, p.deltawin, (p.onlyInSegments ? 1 : 0) | variant};
        if (p.onlyInSegments) {
          segId = -1;
          for (size_t q = 0; q < segComps.size(); q++) if (segComps[q] == s) segId = (int)q;
          if (segId < 0) { segComps.push_back(s); segId = (int)segComps.size() - 1; }
        }
      } else if (s->type == OSM_B200_C_FULLINPUTMEAN) {
        // dspcore/fullinputMean.cpp:484-548 (single EOI loop): all frames are read before EOI, the
        // arithmetic mean is subtracted from every frame at EOI -> same number of frames, and the
        // level is complete only after EOI, so nothing may be chained behind it here
        const auto &p = s->u.fullinputmean;
        if (p.mvn || p.meanNorm != 0 || p.symmSubtract || p.subtractClipToZero || p.specEnorm || p.htkLogEnorm || p.excludeZeros || p.multiLoopMode) {
          err = "cFullinputMean: only plain arithmetic mean subtraction (the defaults) is supported"; return OSM_B200_ERR_UNSUPPORTED;
        }
        if (s != stageComps.back()) { err = "a temporal stage reading a cFullinputMean level is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
        st = Stage{ST_CMS, 0, 0};
      } else {
        const auto &p = s->u.contoursmoother;
        if (p.smaWin < 1 || (p.smaWin & 1) == 0 || p.smaWin > 9) { err = "cContourSmoother.smaWin must be odd, 1..9"; return OSM_B200_ERR_UNSUPPORTED; }
        st = Stage{ST_SMA, (p.smaWin - 1) / 2, p.noZeroSma};
      }
      stages.push_back(st);
      for (auto &f : fields) f.name = name_append_auto(*s, f.name, nullptr);
      stageNo++;
      if (selAbove >= 0 && (long)(stageComps.size() - stageNo) == selAbove) leafSelNames.back() = element_names(fields);
    }
    if (stages.size() > 3) { err = "more than 3 chained temporal stages"; return OSM_B200_ERR_UNSUPPORTED; }
    const size_t firstGroup = d.groups.size();
    int col = d.ops[opIdx].outCol;
    bool open = false;
    for (const auto &f : fields) {
      const bool keep = !(leaf.arraysOnly && f.n == 1);
      if (keep) {
        if (!open) {
          OutGroup g;
          g.srcCol = col; g.n = 0; g.stream = d.ops[opIdx].stream; g.outCol = d.nOut; g.stages = stages;
          if (d.ops[opIdx].kind == SOP_PITCH) { g.lagKind = 1; g.lagOp = opIdx; }
          if (d.ops[opIdx].kind == SOP_JITTER) { g.lagKind = 2; g.lagOp = d.ops[opIdx].jitter.pitchOp; }
          if (d.ops[opIdx].kind == SOP_HARMONICS) { g.lagKind = 1; g.lagOp = d.ops[opIdx].harmonics.pitchOp; }
          if (gateIdx >= 0) {
            const auto &q = gateScopes[gateIdx].c->u.valbasedselector;
            g.gateCol = gateInfo[gateIdx].col;
            g.gateThreshold = (float)q.threshold; g.gateOutVal = (float)q.outputVal;
            g.gateInvert = q.invert != 0; g.gateAllowEqual = q.allowEqual != 0;
            // the gate reads the pitch level: its output ends where that level ends (truncating reader) and lags with it
            if (g.lagKind == 0) g.lagKind = 1;
            if (g.lagOp >= 0 && g.lagOp != gateInfo[gateIdx].pitchOp) { err = "cValbasedSelector: data behind another pitch chain than the selector's"; return OSM_B200_ERR_UNSUPPORTED; }
            g.lagOp = gateInfo[gateIdx].pitchOp;
          }
          g.segId = segId;
          d.groups.push_back(g);
          open = true;
        }
        d.groups.back().n += f.n;
        d.nOut += f.n;
        // element names: name (single element fields) or name[idx + arrNameOffset]
        // (core/dataMemoryLevel.cpp:1158-1169)
        if (f.n == 1) d.names.push_back(f.name);
        else for (int i = 0; i < f.n; i++) {
          snprintf(buf, sizeof buf, "%s[%d]", f.name.c_str(), i + f.arrNameOffset);
          d.names.push_back(buf);
        }
      } else {
        open = false;
      }
      col += f.n;
    }
    leafGroups.push_back({firstGroup, d.groups.size()});
  }
  if (d.ops.empty() || d.groups.empty()) { err = "the output level has no elements (cVectorConcat drops single-element fields unless includeSingleElementFields=1)"; return OSM_B200_ERR_INVALID; }
  // A concat (or multi-level reader) below temporal stages delivers min over its inputs
  // (core/dataReader.cpp:375-380).  Inputs of different frame geometry are supported when they are
  // static levels (no stage below the concat): every group then carries the other streams as limits.
  for (const ConcatCheck &cc : concatChecks) {
    std::vector<int> streams;
    int w0 = -1;
    bool sameW = true;
    for (size_t l = cc.g0; l < cc.g1; l++)
      for (size_t g = leafGroups[l].first; g < leafGroups[l].second; g++) {
        int w = 0;
        for (const auto &st : d.groups[g].stages) w += st.win;
        if (w0 < 0) w0 = w;
        if (w != w0) sameW = false;
        const FrontEnd &fe = d.streams[d.groups[g].stream].fe;
        bool known = false;
        for (int sidx : streams) known = known || (d.streams[sidx].fe.frameSize == fe.frameSize && d.streams[sidx].fe.frameStep == fe.frameStep);
        if (!known) streams.push_back(d.groups[g].stream);
      }
    if (streams.size() <= 1 && sameW) continue;          // nothing truncates
    bool ok = sameW;
    // the stages every leaf carries must all sit above the concat (i.e. the leaves are statics)
    for (size_t l = cc.g0; l < cc.g1 && ok; l++) ok = leaves[l].stages.size() == cc.above;
    if (!ok) { err = "cVectorConcat of unequally long, already smoothed levels below a temporal stage is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
    if (streams.size() > 4) { err = "more than 4 frame geometries below one cVectorConcat"; return OSM_B200_ERR_UNSUPPORTED; }
    for (size_t l = cc.g0; l < cc.g1; l++)
      for (size_t g = leafGroups[l].first; g < leafGroups[l].second; g++)
        for (int sidx : streams) {
          const FrontEnd &a = d.streams[sidx].fe, &b = d.streams[d.groups[g].stream].fe;
          if (a.frameSize == b.frameSize && a.frameStep == b.frameStep) continue;
          if (std::find(d.groups[g].limitStreams.begin(), d.groups[g].limitStreams.end(), sidx) == d.groups[g].limitStreams.end())
            d.groups[g].limitStreams.push_back(sidx);
        }
  }

  // ---- cDataSelector scopes: replace the groups of the leaves below a selector by the selected elements ----
  if (!selScopes.empty()) {
    std::vector<OutGroup> ng;
    std::vector<std::string> nn;
    int outCol = 0;
    for (size_t l = 0; l < leaves.size();) {
      const SelScope *sc = nullptr;
      for (const SelScope &x : selScopes) if (x.l0 == l) sc = &x;
      if (!sc) {
        for (size_t g = leafGroups[l].first; g < leafGroups[l].second; g++) {
          OutGroup x = d.groups[g];
          for (int i = 0; i < x.n; i++) nn.push_back(d.names[x.outCol + i]);
          x.outCol = outCol; outCol += x.n;
          ng.push_back(x);
        }
        l++;
        continue;
      }
      struct El { size_t g; int off; const std::string *name; };
      std::vector<El> els;                                // the elements of the selector's input, in reader order
      for (size_t ll = sc->l0; ll < sc->l1; ll++) {
        size_t e = 0;
        for (size_t g = leafGroups[ll].first; g < leafGroups[ll].second; g++)
          for (int i = 0; i < d.groups[g].n; i++, e++) {
            if (e >= leafSelNames[ll].size()) { err = "internal: cDataSelector element bookkeeping"; return OSM_B200_ERR_INVALID; }
            els.push_back(El{g, i, &leafSelNames[ll][e]});
          }
      }
      const auto &q = sc->c->u.dataselector;
      const std::vector<const osm_b200_component *> &lst = leaves[sc->l0].stages;
      // A selector that reads a cPitchJitter level waits for it: during the reference's first end-of-input pass that
      // level does not advance (lld/pitchJitter.cpp:593), so EVERY element of the selector's output lags like the jitter
      // columns do (oracle/formant_oracle.py:gemaps_lld, pinned on the shipped GeMAPS configurations)
      bool scopeLags = false;
      int scopeLagOp = -1;
      for (const El &e : els) if (d.groups[e.g].lagKind == 2) { scopeLags = true; scopeLagOp = d.groups[e.g].lagOp; }
      if (scopeLags)
        for (size_t ll = sc->l0; ll < sc->l1; ll++)
          if (leaves[ll].stages.size() != sc->above) {
            // pinned only for a selector directly on the per-frame levels (the shipped graphs); a selector over levels
            // that were smoothed first sees another tick order
            err = "cDataSelector reading a cPitchJitter level: temporal stages below the selector are not supported"; return OSM_B200_ERR_UNSUPPORTED;
          }
      long prevG = -1;
      for (int k = 0; k < q.nSelected; k++) {
        for (int k2 = 0; k2 < k; k2++)
          if (!strcmp(q.selected[k], q.selected[k2])) { err = std::string("cDataSelector: element selected twice: ") + q.selected[k]; return OSM_B200_ERR_UNSUPPORTED; }
        const El *hit = nullptr;
        for (const El &e : els) if (*e.name == q.selected[k]) { hit = &e; break; }
        if (!hit) {                                       // core/dataSelector.cpp:330-343 aborts as well
          err = std::string("cDataSelector '") + sc->c->name + "': element '" + q.selected[k] + "' not found in its input levels";
          return OSM_B200_ERR_INVALID;
        }
        OutGroup x = d.groups[hit->g];
        x.srcCol += hit->off; x.n = 1; x.outCol = outCol++;
        if (scopeLags) {
          // every input level of a lagging selector must deliver at least the rows of the pitch level (same step, frames
          // not longer): the selector's output then has the pitch level's length and the limits of the truncating reader
          // are redundant.  A per-frame level routed through it (e.g. the formants of the 20 ms frames) lags like the rest.
          const FrontEnd &fp = d.streams[d.ops[scopeLagOp].stream].fe;
          std::vector<int> chk = x.limitStreams;
          chk.push_back(x.stream);
          for (int sidx : chk) {
            const FrontEnd &fx = d.streams[sidx].fe;
            if (fx.frameStep != fp.frameStep || fx.frameSize > fp.frameSize) {
              err = "cDataSelector reading a cPitchJitter level together with a level of another frame step / longer frames is not supported"; return OSM_B200_ERR_UNSUPPORTED;
            }
          }
          x.limitStreams.clear();
          x.lagKind = 2; x.lagOp = scopeLagOp; x.stream = d.ops[scopeLagOp].stream;
        }
        if (prevG == (long)hit->g && !ng.empty() && ng.back().srcCol + ng.back().n == x.srcCol) ng.back().n++;
        else ng.push_back(x);
        prevG = (long)hit->g;
        std::string nm = q.newNames[k][0] ? std::string(q.newNames[k])                                  // :344-356
                         : (sc->c->nameAppend[0] ? std::string(q.selected[k]) + "_" + sc->c->nameAppend : std::string(q.selected[k]));
        for (size_t a = lst.size() - sc->above; a < lst.size(); a++) nm = name_append_auto(*lst[a], nm, nullptr);
        nn.push_back(nm);
      }
      l = sc->l1;
    }
    d.groups.swap(ng);
    d.names.swap(nn);
    d.nOut = outCol;
  }

  // ---- gated groups: every data level must deliver at least the rows of the pitch level (same step, frames not longer), the gate's
  // output then has the pitch level's length ----
  for (OutGroup &g : d.groups) {
    if (g.gateCol < 0) continue;
    const int ps = d.ops[g.lagOp].stream;
    const FrontEnd &fp = d.streams[ps].fe;
    std::vector<int> chk = g.limitStreams;
    chk.push_back(g.stream);
    for (int sidx : chk) {
      const FrontEnd &fx = d.streams[sidx].fe;
      if (fx.frameStep != fp.frameStep || fx.frameSize > fp.frameSize) { err = "cValbasedSelector: a data level of another frame step / longer frames than the pitch level is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
    }
    g.limitStreams.clear();
    g.stream = ps;
    if (g.stages.empty()) { err = "a cValbasedSelector level as output level (no cContourSmoother behind it) is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
  }

  // ---- groups behind a Viterbi-smoothed pitch level (seq_post_kernel) ----
  // Supported shapes (the ones the shipped feature sets use): [cContourSmoother(3)] and
  // [cContourSmoother(3), cDeltaRegression(onlyInSegments)], no truncating concat in between.
  for (const OutGroup &g : d.groups) {
    const bool seg = g.segId >= 0;
    if (!seg && g.lagKind == 0) continue;
    if (g.lagKind == 0) { err = "cDeltaRegression.onlyInSegments=1 is only supported behind the SHS pitch chain (cPitchSmootherViterbi / cPitchJitter levels)"; return OSM_B200_ERR_UNSUPPORTED; }
    const size_t ns = g.stages.size();
    bool ok = ns >= 1 && ns <= 2 && g.stages[0].kind == ST_SMA && g.stages[0].win == 1;
    if (ok && ns == 2) ok = g.stages[1].kind == ST_DELTA && (g.stages[1].flags & 1) && g.stages[1].win >= 1 && g.stages[1].win <= 4;
    if (!ok) { err = "levels behind cPitchSmootherViterbi / cPitchJitter support cContourSmoother(smaWin=3) optionally followed by cDeltaRegression(onlyInSegments=1) only"; return OSM_B200_ERR_UNSUPPORTED; }
    if (!g.limitStreams.empty()) { err = "a truncating concat below the temporal stages of a pitch level is not supported"; return OSM_B200_ERR_UNSUPPORTED; }
  }

  // ---- execution strategy per stream ----
  // A stream with exactly one band op (MFCC / PLP) and no other spectral consumer evaluates it
  // inside lld_kernel; any other spectral consumer reads the magnitude level from HBM.
  for (size_t s = 0; s < d.streams.size(); s++) {
    int nSpec = 0;
    d.streams[s].bandOps.clear();
    for (size_t o = 0; o < d.ops.size(); o++) {
      if (d.ops[o].stream != (int)s) continue;
      if (d.ops[o].kind == SOP_MFCC || d.ops[o].kind == SOP_PLP) d.streams[s].bandOps.push_back((int)o);
      if (d.ops[o].kind == SOP_SPECTRAL || d.ops[o].kind == SOP_PITCHACF || d.ops[o].kind == SOP_MAG || d.ops[o].kind == SOP_PITCH || d.ops[o].kind == SOP_HARMONICS) nSpec++;
    }
    const int nBand = (int)d.streams[s].bandOps.size();
    d.streams[s].fusedOp = nBand ? d.streams[s].bandOps[0] : -1;
    d.streams[s].dumpMag = nSpec > 0;
    if ((nBand || nSpec) && (d.streams[s].fe.nfft < 64 || d.streams[s].fe.nfft > 4096)) { err = "FFT size out of the supported range"; return OSM_B200_ERR_UNSUPPORTED; }
  }
  return OSM_B200_OK;
}

}  // namespace osm
