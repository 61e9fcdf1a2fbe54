"""CPU tests of the drop-in boundary: libglcuda.so loads without a GPU, exports every symbol that
include/gl_cuda.h declares, and fails loudly (no CPU fallback) when asked to compute."""
import os
import re

import pytest

import ipgl_b200 as gl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gl_cuda.h")).read()
    return sorted(set(re.findall(r"GL_API\s+[\w\s\*]+?\b(gl_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 35
    L = gl.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(gl.EXPORTS) == names          # the ctypes binding covers the whole header


def test_version_and_defaults():
    assert gl.lib().gl_version() >= 100
    p = gl.default_params()
    # hpc/affinity.c:117-118, hpc/display.c:73, hpc/utils.c:721 (MatPow no-op)
    assert (p.h_loc, p.h_val, p.gain, p.power) == (40.0, 30.0, 3.0, 1.0)
    assert p.affinity_kind == 0 and p.sampling_random == 0 and p.num_eigvals == -1 and p.gram_schmidt == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gl.GLError) as e:
        gl.Context(0)
    assert e.value.status == gl.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "image-processing-graph-laplacian_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".c", ".h", ".py", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "oracle_c" not in text and "from oracle" not in text, f


def test_header_is_plain_c(tmp_path):
    """include/gl_cuda.h is the drop-in boundary: it must compile as C (no C++ or torch types) on its own."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "gl_cuda.h"\nint main(void) { gl_params p; gl_default_params(&p); return (int)sizeof(gl_mat_info) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_host_defines_every_reference_entry_point():
    """Every stage entry point the reference's hpc/*.h declare is declared in hpc/hpc_api.h and defined by the C host
    (checked on the linked binary's symbol table; static helpers do not count)."""
    import subprocess
    hpc = os.path.join(ROOT, "image-processing-graph-laplacian_b200", "hpc")
    api = open(os.path.join(hpc, "hpc_api.h")).read()
    want = ["Sampling", "ComputeAffinityMatrices", "ComputeEntireAffinityMatrix", "ComputeLaplacianMatrix", "ComputeEntireLaplacianMatrix",
            "EigendecompositionLargest", "EigendecompositionSmallest", "InversePowerIteration", "Nystroem", "OrthonormaliseVecs",
            "NormaliseVecs", "WriteVec", "WriteDiagMat", "WriteMatCol", "WritePngMatCol", "ComputeResultFromLaplacian",
            "ComputeResultFromEntireLaplacian", "Permutation", "MatRowSum", "VecMean", "InverseDiagMat", "MatPow", "DiagMat2Vec",
            "num2x", "num2y", "xy2num", "read_png", "write_png"]
    for name in want:
        assert re.search(r"\b%s\s*\(" % name, api), name
    binary = os.path.join(hpc, "image_processing")
    if not os.path.exists(binary):
        subprocess.check_call(["make", "-C", hpc])
    syms = subprocess.run(["nm", "--defined-only", binary], capture_output=True, text=True).stdout
    defined = set(line.split()[-1] for line in syms.splitlines() if line.strip())
    missing = [n for n in want if n not in defined]
    assert not missing, missing
