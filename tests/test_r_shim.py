"""The R `.Call` shim (r-package/src/shim.c) is compiled -- with -Wall -Wextra -Werror -- and RUN against a minimal
stand-in for R's C API (tests/r_stub/), because R itself is absent from this image and from the GPU box.
CPU: it builds, links against libscgbm.so, and a fit without a device raises R's error with the library's message.
GPU: gbm.sc -> get.se -> CCI perturbation through the shim give bit for bit what the ctypes mirror gives, for dense
and dgCMatrix input, with and without W (the NULL-W path of round 1's review), and over several GPUs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from scgbm_b200 import _lib as L
    L.load()                                          # libscgbm.so must exist (no fallback)
    d = tmp_path_factory.mktemp("rshim")
    cc = ["gcc", "-std=c11", "-O1", "-Wall", "-I", STUB, "-I", os.path.join(ROOT, "include")]
    objs = []
    for src, extra in ((os.path.join(ROOT, "r-package", "src", "shim.c"), ["-Wextra", "-Werror", "-Wno-cast-function-type"]),
                       (os.path.join(STUB, "rstub.c"), []), (os.path.join(STUB, "harness.c"), [])):
        o = os.path.join(d, os.path.basename(src) + ".o")
        subprocess.check_call(cc + extra + ["-c", src, "-o", o])
        objs.append(o)
    exe = os.path.join(d, "harness")
    libdir = os.path.join(ROOT, "scgbm_b200")
    subprocess.check_call(["gcc", "-o", exe] + objs + ["-L", libdir, "-lscgbm", "-Wl,-rpath," + libdir])
    return exe, str(d)


def _write_input(path, Y, M, max_iter, sparse, return_W, ngpu=1):
    I, J = Y.shape
    with open(path, "wb") as f:
        np.array([I, J, M, max_iter, int(sparse), int(return_W), ngpu], dtype=np.int32).tofile(f)
        np.asfortranarray(Y.astype(np.int32)).T.copy().tofile(f)        # column-major bytes


def _read_output(path, names):
    buf = np.fromfile(path, dtype=np.float64)
    n = int(buf[0]); pos = 1
    assert n == len(names)
    vals = []
    for _ in range(n + 3):
        ln = int(buf[pos]); vals.append(buf[pos + 1: pos + 1 + ln]); pos += 1 + ln
    assert pos == len(buf)
    return dict(zip(names, vals[:n])), vals[n], vals[n + 1], vals[n + 2]


def test_shim_compiles_and_raises_the_library_error_without_a_device(harness):
    from scgbm_b200 import _lib as L
    exe, d = harness
    if L.load().scgbm_device_count() > 0:
        pytest.skip("a GPU is present")
    Y = np.random.default_rng(0).poisson(2.0, (40, 50))
    _write_input(os.path.join(d, "in.bin"), Y, 3, 10, False, True)
    r = subprocess.run([exe, "fit", os.path.join(d, "in.bin"), os.path.join(d, "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3                                    # error() -> R's top level
    assert "no CUDA device available" in r.stderr


@pytest.mark.parametrize("sparse,return_W", [(False, True), (True, False)])
def test_shim_list_assembly_without_a_device(tmp_path, sparse, return_W):
    """The shim's marshalling executed on the CPU: harness + shim + R stand-in linked against a do-nothing stand-in
    for libscgbm (tests/r_stub/fake_scgbm.c).  Catches what compiling alone cannot -- round 2 found the returned
    list one element too short (a write past the end of `out`) this way."""
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = os.path.join(tmp_path, "harness_fake")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-I", STUB, "-I", os.path.join(ROOT, "include"),
                           os.path.join(STUB, "harness.c"), os.path.join(ROOT, "r-package", "src", "shim.c"),
                           os.path.join(STUB, "rstub.c"), os.path.join(STUB, "fake_scgbm.c"), "-o", exe])
    Y = np.random.default_rng(0).poisson(2.0, (6, 7))
    _write_input(os.path.join(tmp_path, "in.bin"), Y, 2, 10, sparse, return_W)
    r = subprocess.run([exe, "fit", os.path.join(tmp_path, "in.bin"), os.path.join(tmp_path, "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    names = [ln for ln in r.stdout.strip().split("\n") if not ln.startswith(("Iteration", "warnings="))]
    base = ["V", "D", "U", "loadings", "alpha", "beta", "M"] + (["I", "J"] if return_W else []) + ["obj", "scores"]
    assert names == (["W"] if return_W else []) + base + ([] if return_W else ["Xu", "Xv", "x.rank"])
    out, se_s, se_l, pert = _read_output(os.path.join(tmp_path, "out.bin"), names)
    assert out["D"][0] == 42.0 and out["obj"].tolist() == [-1.0, -0.5, -0.25] and out["M"][0] == 2
    assert out["V"].shape == (7 * 2,) and out["U"].shape == (6 * 2,) and out["alpha"].shape == (6,)
    assert se_s[0] == (1.0 if return_W else 2.0)          # get.se took the W path / the kept-factor path
    assert pert[0] == 3.0 and pert.shape == (7 * 2 * 3,)
    if not return_W:
        assert out["x.rank"][0] == 2
        r = subprocess.run([exe, "badcall", os.path.join(tmp_path, "in.bin"), os.path.join(tmp_path, "o2.bin")], capture_output=True, text=True)
        assert r.returncode == 3 and "get.se needs out$W" in r.stderr


def test_shim_sources_declare_every_registered_routine():
    """the .Call names used by r-package/R/*.R are exactly the routines registered in the shim"""
    import re
    shim = open(os.path.join(ROOT, "r-package", "src", "shim.c")).read()
    registered = set(re.findall(r'\{"(scgbm_\w+_R)"', shim))
    used = set()
    for f in os.listdir(os.path.join(ROOT, "r-package", "R")):
        used |= set(re.findall(r'\.Call\("(scgbm_\w+_R)"', open(os.path.join(ROOT, "r-package", "R", f)).read()))
    assert used and used <= registered, used - registered
    assert registered - used <= {"scgbm_device_count_R"}


@pytest.mark.gpu
@pytest.mark.parametrize("sparse,return_W,ngpu", [(False, True, 1), (True, True, 1), (False, False, 1), (True, False, 2)])
def test_shim_fit_get_se_and_cci_equal_the_ctypes_mirror(harness, sparse, return_W, ngpu):
    from scgbm_b200 import _lib as L, gbm_sc, get_se, cci_perturb
    from oracle.scgbm_oracle import sim_data
    if L.load().scgbm_device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    exe, d = harness
    Y = sim_data(220, 340, 4, seed=13)["Y"]
    M, max_iter = 4, 40
    _write_input(os.path.join(d, "in.bin"), Y, M, max_iter, sparse, return_W, ngpu)
    r = subprocess.run([exe, "fit", os.path.join(d, "in.bin"), os.path.join(d, "out.bin")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    # (NCCL prints its version banner on stdout when the in-process multi-GPU driver creates the communicators)
    names = [ln for ln in lines if not ln.startswith(("Iteration", "warnings=", "NCCL version"))]
    base = ["V", "D", "U", "loadings", "alpha", "beta", "M"] + (["I", "J"] if return_W else []) + ["obj", "scores"]
    assert names == (["W"] if return_W else []) + base + ([] if return_W else ["Xu", "Xv", "x.rank"])   # R/fit.R:189-204,310
    out, se_s, se_l, pert = _read_output(os.path.join(d, "out.bin"), names)
    if ngpu > 1:     # the in-process multi-GPU driver is exercised in a child (tests/test_gpu_multi.py explains why);
        ngpu_ref = 1  # here the shim's ngpu run is compared with the one-GPU mirror to the parity tolerances
    else:
        ngpu_ref = ngpu
    ref = gbm_sc(Y, M, max_iter=max_iter, return_W=return_W, keep_factors=not return_W, ngpu=ngpu_ref)
    ref = get_se(ref, ngpu=ngpu_ref)
    I, J = Y.shape
    if ngpu > 1:
        from tests.util import principal_angle
        assert principal_angle(out["U"].reshape((I, M), order="F"), ref["U"]) < 1e-4
        np.testing.assert_allclose(out["obj"], ref["obj"][: len(out["obj"])], rtol=1e-6)
        np.testing.assert_allclose(out["alpha"], ref["alpha"].ravel(order="F"), atol=1e-5)
        np.testing.assert_allclose(se_s.reshape((J, M), order="F"), ref["se_scores"], rtol=1e-3)
        return
    # the same library, the same deterministic kernels: bit for bit
    np.testing.assert_array_equal(out["U"].reshape((I, M), order="F"), ref["U"])
    np.testing.assert_array_equal(out["V"].reshape((J, M), order="F"), ref["V"])
    np.testing.assert_array_equal(out["obj"], ref["obj"])
    np.testing.assert_array_equal(out["alpha"], ref["alpha"].ravel(order="F"))
    np.testing.assert_array_equal(out["loadings"].reshape((I, M), order="F"), ref["loadings"])
    np.testing.assert_array_equal(se_s.reshape((J, M), order="F"), ref["se_scores"])
    np.testing.assert_array_equal(se_l.reshape((I, M), order="F"), ref["se_loadings"])
    if return_W:
        np.testing.assert_array_equal(out["W"].reshape((I, J), order="F"), ref["W"])
        assert out["I"][0] == I and out["J"][0] == J
    np.testing.assert_array_equal(pert.reshape((J, M, 3), order="F"), cci_perturb(ref, reps=3, seed=21))
    # the progress lines of R/fit.R:175, one per accepted iteration
    assert sum(ln.startswith("Iteration:") for ln in lines) == len(ref["obj"])
    # get.se without W and without factors: R's error, not a crash (round-1 review, INTEGER(NULL))
    if not return_W:
        r = subprocess.run([exe, "badcall", os.path.join(d, "in.bin"), os.path.join(d, "out2.bin")], capture_output=True, text=True)
        assert r.returncode == 3 and "get.se needs out$W" in r.stderr
