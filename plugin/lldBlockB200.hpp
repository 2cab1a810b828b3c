/*
 * cLldBlockB200 -- openSMILE plugin component: block execution of an LLD sub-graph on a B200 through libosm_b200.so.
 *
 * Compiled against the reference's own headers (src/include) and loaded by the reference's plugin loader
 * (src/core/componentManager.cpp:212-425, entry point plugindev/pluginMain.cpp:43-58).  It is a cDataProcessor
 * (src/include/core/dataProcessor.hpp:26-142) like every LLD component: it reads the `wave` level the stock
 * cWaveSource / cExternalAudioSource writes, and writes the level the stock sinks (cHtkSink, cCsvSink, cArffSink,
 * cExternalSink, cFunctionals ...) read.  Everything between those two levels -- the chain
 * cFramer -> ... -> cVectorConcat of a shipped .conf -- is replaced by ONE batched CUDA plan run at end of input.
 *
 * Configuration (in addition to cDataProcessor's reader / writer / buffersize fields):
 *   graphConf      the configuration file whose LLD sub-graph is executed (a shipped .conf, read unchanged)
 *   captureTo      the level of that graph to produce (default: the level its active sinks read)
 *   graphOption[]  "name=value" command-line options of the graph file (its \cm[...] fields, e.g. lldcsvoutput=x)
 *   device         CUDA device (default 0).  There is no CPU path: without a device the component fails to configure.
 */
#ifndef OSM_B200_PLUGIN_LLDBLOCK_HPP
#define OSM_B200_PLUGIN_LLDBLOCK_HPP

#include <core/smileCommon.hpp>
#include <core/dataProcessor.hpp>

#include <string>
#include <vector>

#include "osm_b200_host.h"

#define COMPONENT_DESCRIPTION_CLLDBLOCKB200 "Block execution of an LLD sub-graph (framing ... concatenation of a feature configuration file) on an NVIDIA B200 through libosm_b200.so: gathers the wave level until the end of input, runs one batched CUDA plan and writes the resulting LLD rows."
#define COMPONENT_NAME_CLLDBLOCKB200 "cLldBlockB200"

class cLldBlockB200 : public cDataProcessor {
 private:
  osm_b200_session *session_;
  osm_b200_plan *plan_;            // owned by the session
  const char *graphConf_, *captureTo_;
  int device_;
  std::vector<std::string> optNames_, optValues_;
  double sampleRate_;
  long nEl_;
  std::vector<float> wave_;        // the gathered wave level (one float per sample frame, mono)
  std::vector<float> rows_;        // LLD rows of the block
  long nRows_, nTimeRows_, emitted_;
  bool ran_;

  void openSession();
  void runBlock();
  static bool recoverPcm(const std::vector<float> &wave, std::vector<int16_t> &pcm, int &nCarrier);

 protected:
  SMILECOMPONENT_STATIC_DECL_PR

  virtual void myFetchConfig() override;
  virtual int configureReader(const sDmLevelConfig &c) override;
  virtual int configureWriter(sDmLevelConfig &c) override;
  virtual int setupNewNames(long nEl) override;
  virtual eTickResult myTick(long long t) override;

 public:
  SMILECOMPONENT_STATIC_DECL

  cLldBlockB200(const char *_name);
  virtual ~cLldBlockB200();
};

#endif
