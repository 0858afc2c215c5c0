"""oracle/patch_model_manager.py -- TEST INFRASTRUCTURE ONLY (build step of oracle/_ref/libdropin_ref.so).

Applies the registration patch that INTEGRATION.md / rvtests_b200/host/ModelB200.h describe to a SCRATCH copy of the
reference's src/ModelManager.cpp (read from /root/reference, written under /tmp; never into the repository):
one #include and one macro line in front of the first model of the "burden", "kernel" and "meta" branches."""
import sys

src, dst = sys.argv[1], sys.argv[2]
s = open(src).read()
s = s.replace('#include "src/Model.h"', '#include "src/Model.h"\n#include "ModelB200.h"', 1)
for branch, first, macro in (('modelType == "burden"', 'if (modelName == "cmc") {', "RVT_B200_BURDEN_MODELS(modelName, parser, model)"),
                             ('modelType == "kernel"', 'if (modelName == "skat") {', "RVT_B200_KERNEL_MODELS(modelName, parser, model)"),
                             ('modelType == "meta"', 'if (modelName == "score") {', "RVT_B200_META_MODELS(modelName, parser, model)")):
    i = s.index(branch)
    j = s.index(first, i)
    s = s[:j] + macro + "\n    " + s[j:]
open(dst, "w").write(s)
print("patched", src, "->", dst)
