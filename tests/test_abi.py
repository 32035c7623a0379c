"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/scgbm.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "scgbm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scgbm_[a-z0-9_]+)\s*\(", src)) - {"scgbm_progress_cb"})


def test_header_symbols_are_exported():
    from scgbm_b200 import _lib as L
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/scgbm.h but not exported"
    assert sorted(L.EXPORTS) == names
    assert lib.scgbm_version() == 200


def test_struct_layouts_match_header():
    """ctypes mirrors must have the C sizes (checked against a tiny C program when gcc is there)."""
    import shutil
    import subprocess
    import tempfile
    from scgbm_b200 import _lib as L
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    prog = r'''
#include <stdio.h>
#include "scgbm.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(scgbm_simdata_args), sizeof(scgbm_simdata_out), sizeof(scgbm_cci_args), sizeof(scgbm_profile), sizeof(scgbm_fit_args),
         sizeof(scgbm_fit_out), sizeof(scgbm_se_args), sizeof(scgbm_se_out), sizeof(scgbm_proj_args),
         sizeof(scgbm_proj_out), sizeof(scgbm_sim_args), sizeof(scgbm_tsvd_args), sizeof(scgbm_tsvd_out),
         sizeof(scgbm_eval_args), sizeof(scgbm_eval_out), sizeof(scgbm_prod_args), sizeof(scgbm_prod_out));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    mine = [C.sizeof(t) for t in (L.SimDataArgs, L.SimDataOut, L.CciArgs, L.Profile, L.FitArgs, L.FitOut, L.SeArgs, L.SeOut, L.ProjArgs, L.ProjOut,
                                  L.SimArgs, L.TsvdArgs, L.TsvdOut, L.EvalArgs, L.EvalOut, L.ProdArgs, L.ProdOut)]
    assert mine == sizes


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry fails loudly (status 2)."""
    from scgbm_b200 import _lib as L, gbm_sc
    lib = L.load()
    if lib.scgbm_device_count() > 0:
        pytest.skip("a GPU is present")
    Y = np.ones((8, 9), dtype=np.int32)
    with pytest.raises(L.ScgbmError, match="no CUDA device") as ei:
        gbm_sc(Y, 2)
    assert ei.value.status == 2
    p = C.c_void_p()
    assert lib.scgbm_dev_alloc(C.byref(p), 16, -1) == 2


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "scgbm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions oracle"
