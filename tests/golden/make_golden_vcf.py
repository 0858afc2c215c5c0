"""Regenerates tests/golden/vcf_golden.json.  Run in the BUILD container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_vcf.py
Synthetic VCF records (tests/test_vcf_pack.py::_random_record, plus a few hand-written lines with the GT quirks) and the
genotypes the REFERENCE's own record parser -- libVcf compiled unmodified into oracle/_ref/libvcf_ref.so -- returned."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
from test_vcf_pack import _header, _random_record  # noqa: E402

O.build()
assert O.ref_vcf() is not None, "oracle/_ref/libvcf_ref.so not built"
n = 11
hdr = _header(n)
rng = np.random.default_rng(20260925)
recs = [_random_record(rng, n, 100 + 7 * k) for k in range(40)]
site = ["1", "1000", "q", "G", "A,C", "69", "PASS", "NS=3"]
recs.append("\t".join(site + ["GT:GD:GQ", "./0:0:29", "2/1:45:100"] + ["0/0:1:32"] * (n - 2)))
recs.append("\t".join(site[:1] + ["1001"] + site[2:] + ["GD:GQ"] + ["3:4"] * n))                    # no GT key at all
recs.append("\t".join(site[:1] + ["1002"] + site[2:] + ["GQ:GT", "5", "5:1/1"] + ["7:0|1"] * (n - 2)))  # truncated column
out = dict(header=hdr, records=recs, genotypes=[], sites=[], gt_strings={})
for r in recs:
    chrom, pos, g = O.ref_vcf_genotypes(hdr, r)
    out["genotypes"].append([int(x) for x in g])
    out["sites"].append(f"{chrom}:{pos}")
for s in ("0/1", "1/0", "1|1", "0", "1", "2", ".", "./.", "0/2", "2/1", "0/", "1/.", "0/-", "0/1/1", "", "0|0", "1/1", "a/0", "0/9"):
    out["gt_strings"][s] = int(O.ref_vcf().ref_vcf_gt(s.encode(), len(s)))
json.dump(out, open(os.path.join(HERE, "vcf_golden.json"), "w"), indent=0)
print(len(recs), "records;", out["gt_strings"])
