// ModelB200.h -- the file a maintainer drops into rvtests' src/ to put the B200 engine behind the existing plugin
// surface.  It instantiates the adapters of rvt_fitters.h with the reference's OWN types, as true ModelFitter
// subclasses (src/ModelFitter.h:17-75), and provides the two registration macros for ModelManager::create
// (src/ModelManager.cpp:99-103 "burden", :168-187 "kernel"):
//
//     #include "ModelB200.h"
//     ...
//     } else if (modelType == "burden") {
//       RVT_B200_BURDEN_MODELS(modelName, parser, model)        // <- added
//       if (modelName == "cmc") {
//     ...
//     } else if (modelType == "kernel") {
//       RVT_B200_KERNEL_MODELS(modelName, parser, model)        // <- added
//       if (modelName == "skat") {
//     ...
//     } else if (modelType == "meta") {
//       RVT_B200_META_MODELS(modelName, parser, model)          // <- added
//       if (modelName == "score") {
//
// The B200 models are created when the environment variable RVTESTS_B200 is set to a non-empty value other than "0"
// (a maintainer would add a command-line flag next to the other ones in src/Main.cpp); otherwise the stock models run.
// Everything downstream -- setParameter/setPrefix/setBinaryOutcome, one FileWriter per model, the gene loop
// reset()/fit(&dc)/writeOutput() (src/Main.cpp:1249-1253), writeFootnote + delete at ModelManager::close -- is the
// reference's code, unchanged.  oracle/Makefile applies exactly this patch to a scratch copy of src/ModelManager.cpp,
// compiles it against the reference tree, and tests/test_gpu_dropin.py runs the reference's own fitters and the B200
// fitters through it, side by side, on a B200.
#ifndef RVT_MODEL_B200_H_
#define RVT_MODEL_B200_H_

#include <stdlib.h>

#include "base/IO.h"               // FileWriter
#include "src/DataConsolidator.h"
#include "src/ModelFitter.h"
#include "src/ModelParser.h"
#include "src/Result.h"

#include "rvt_fitters.h"
#include "rvt_meta_fitters.h"

typedef rvtb200::SkatTestB200<DataConsolidator, FileWriter, Result, ModelFitter> SkatTestB200;
typedef rvtb200::SkatOTestB200<DataConsolidator, FileWriter, Result, ModelFitter> SkatOTestB200;
typedef rvtb200::CMCTestB200<DataConsolidator, FileWriter, Result, ModelFitter> CMCTestB200;
typedef rvtb200::ZegginiTestB200<DataConsolidator, FileWriter, Result, ModelFitter> ZegginiTestB200;
typedef rvtb200::MetaScoreTestB200<DataConsolidator, FileWriter, Result, ModelFitter> MetaScoreTestB200;
typedef rvtb200::MetaCovTestB200<DataConsolidator, FileWriter, Result, ModelFitter> MetaCovTestB200;

#include "src/Summary.h"
extern SummaryHeader* g_SummaryHeader;   // src/Main.cpp / src/Model.cpp

namespace rvtb200 {
// the reference's own summary block and covariate labels, read when the first line is written
struct RefSummaryHook : SummaryHook<FileWriter> {
  virtual void outputHeader(FileWriter* fp) {
    if (g_SummaryHeader) g_SummaryHeader->outputHeader(fp);
  }
  virtual const std::vector<std::string>& getCovLabel() const {
    static const std::vector<std::string> none;
    return g_SummaryHeader ? g_SummaryHeader->getCovLabel() : none;
  }
  static RefSummaryHook* instance() {
    static RefSummaryHook h;
    return &h;
  }
};
inline ::MetaScoreTestB200* newMetaScore(bool se) {
  ::MetaScoreTestB200* m = new ::MetaScoreTestB200(se);
  m->setSummaryHeader(RefSummaryHook::instance());
  return m;
}
inline ::MetaCovTestB200* newMetaCov(int windowSize) {
  ::MetaCovTestB200* m = new ::MetaCovTestB200(windowSize);
  m->setSummaryHeader(RefSummaryHook::instance());
  return m;
}
inline bool enabled() {
  const char* e = getenv("RVTESTS_B200");
  return e && e[0] && !(e[0] == '0' && e[1] == 0);
}
}  // namespace rvtb200

#define RVT_B200_BURDEN_MODELS(modelName, parser, model)        \
  if (rvtb200::enabled() && (modelName) == "cmc") {             \
    (model).push_back(new CMCTestB200);                         \
  } else if (rvtb200::enabled() && (modelName) == "zeggini") {  \
    (model).push_back(new ZegginiTestB200);                     \
  } else

#define RVT_B200_KERNEL_MODELS(modelName, parser, model)                                   \
  if (rvtb200::enabled() && (modelName) == "skat") {                                       \
    int nPermB200 = 10000;                                                                 \
    double alphaB200 = 0.05, beta1B200 = 1.0, beta2B200 = 25.0;                            \
    (parser).assign("nPerm", &nPermB200, 10000).assign("alpha", &alphaB200, 0.05)          \
        .assign("beta1", &beta1B200, 1.0).assign("beta2", &beta2B200, 25.0);               \
    (model).push_back(new SkatTestB200(nPermB200, alphaB200, beta1B200, beta2B200));       \
  } else if (rvtb200::enabled() && (modelName) == "skato") {                               \
    double beta1B200 = 1.0, beta2B200 = 25.0;                                              \
    (parser).assign("beta1", &beta1B200, 1.0).assign("beta2", &beta2B200, 25.0);           \
    (model).push_back(new SkatOTestB200(beta1B200, beta2B200));                            \
  } else

// --meta score[se],cov[windowSize=..]  (src/ModelManager.cpp:208-236), quantitative or binary trait, unrelated samples
#define RVT_B200_META_MODELS(modelName, parser, model)                               \
  if (rvtb200::enabled() && (modelName) == "score") {                                \
    (model).push_back(rvtb200::newMetaScore((parser).hasTag("se")));                 \
  } else if (rvtb200::enabled() && (modelName) == "cov") {                           \
    int windowSizeB200 = 1000000;                                                    \
    (parser).assign("windowSize", &windowSizeB200, 1000000);                         \
    (model).push_back(rvtb200::newMetaCov(windowSizeB200));                          \
  } else

#endif  // RVT_MODEL_B200_H_
