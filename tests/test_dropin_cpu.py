"""CPU: the literal drop-in build (oracle/_ref/libdropin_ref.so = the reference's own ModelManager.cpp with the registration
patch of rvtests_b200/host/ModelB200.h + Main.cpp's gene loop, oracle/ref_dropin_shim.cpp).  Without a GPU this checks
(1) that the patched ModelManager's stock path prints exactly what the unpatched reference model layer prints, and (2) that
the B200 adapters are created by name, run through the ModelFitter interface, and -- there being no CPU fallback -- print
the reference's NA columns under the reference's own header.  The numbers are compared on the GPU (tests/test_gpu_dropin.py)."""
import numpy as np
import pytest

from test_gpu_dropin import _genes


def test_patched_model_manager_stock_path_and_na_without_gpu(oracle, tmp_path):
    O = oracle
    if O.ref_dropin() is None or O.ref_model() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    genes, X, y = _genes(O, 301, 600, 2)
    ref = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "ref"), use_b200=False)
    stock = O.ref_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "stock"), n_perm=0)
    for model in ("Skat", "SkatO", "CMC", "Zeggini"):
        assert stock[model][1] == ref[model][1] and stock[model][2] == ref[model][2], model
    import conftest
    if conftest._cuda_device_present():
        return
    b2 = O.dropin_run_gene_models(genes, X[:, 1:], y, str(tmp_path / "b200"), use_b200=True)
    for model in ("Skat", "SkatO", "CMC", "Zeggini"):
        assert b2[model][1] == ref[model][1], model                      # same header
        assert len(b2[model][2]) == len(ref[model][2])
        for lr, lb in zip(ref[model][2], b2[model][2]):
            assert lb[:5] == lr[:5] and all(v == "NA" for v in lb[5:]), (model, lb)
