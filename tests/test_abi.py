"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/sgp.h declares,
the ctypes table covers them all, and -- there being no CPU fallback -- creating a context without a GPU fails loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from gaussianprocessnode_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("sgp_create", "sgp_destroy", "sgp_set_kernel", "sgp_set_inducing", "sgp_set_data", "sgp_sweep_psi",
                 "sgp_sweep_psi_uncertain", "sgp_kuu_factor", "sgp_kuu_solve", "sgp_posterior_v", "sgp_w_terms",
                 "sgp_predict_mean", "sgp_comm_unique_id", "sgp_comm_init", "sgp_sweep_psi_host", "sgp_sweep_psi_host_packed", "sgp_fetch_psi2_packed", "sgp_theta_objective",
                 "sgp_in_logmessage", "sgp_uncertain_node_terms", "sgp_sweep_timed_flushed"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from gaussianprocessnode_b200 import _lib
    for name in _declared():
        assert hasattr(lib, name), "libsgp.so does not export %s" % name
        assert name in _lib.SIGNATURES, "ctypes table misses %s" % name
    assert set(_lib.SIGNATURES) == set(_declared())
    assert b"sm_100a" in lib.sgp_version()


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "sgp.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S).lower()
    assert "at::" not in src and "cudaStream_t" not in src


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.sgp_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    from gaussianprocessnode_b200 import SGPContext, SGPError
    with pytest.raises(SGPError):
        SGPContext(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gaussianprocessnode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def build_c_client(tmp_path):
    """examples/c_client.c with gcc as plain C11 against include/sgp.h and libsgp.so; returns the binary's path."""
    import subprocess
    pkg = os.path.join(ROOT, "gaussianprocessnode_b200")
    exe = str(tmp_path / "c_client")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_client.c"), "-o", exe,
                    "-L" + pkg, "-lsgp", "-lm", "-Wl,-rpath," + pkg, "-Wl,--unresolved-symbols=ignore-in-shared-libs"], check=True, timeout=120)
    return exe


def test_plain_c_client_compiles_links_and_fails_loudly_without_a_gpu(lib, tmp_path):
    # the header is valid C (not only C++), the entry points link from C, and without a device the first call fails with SGP_ERR_CUDA
    import subprocess
    import torch
    exe = build_c_client(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU test runs the client")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "sgp_create" in r.stderr and "-> -2" in r.stderr
