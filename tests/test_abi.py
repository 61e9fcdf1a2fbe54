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
