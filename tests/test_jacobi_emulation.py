"""The schedule of csrc/eigen_jacobi.cu emulated in numpy (tools/emulate_jacobi_cross.py): rotating only the cross pairs of a panel
pair outside the panels' home visit (what the kernel does while a sweep starts above 4 tol) must converge in as many outer sweeps as the
full cyclic inner sweep, to the same eigenvalues.  CPU only: this pins the algorithmic claim the kernel's comment makes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("tag", ["barbara_uniform256", "lion_photometric_h10", "cat_small_random50"])
def test_cross_only_inner_sweeps_converge_like_full_ones(tag):
    import emulate_jacobi_cross as em
    from oracle import oracle_np as o
    g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    img = g["image"]
    r = o.run_pipeline(img, g["sample_indices"], kind=str(g["kind"]), h_loc=float(g["h_loc"]), h_val=float(g["h_val"]))
    L = r["L_A"]
    s_full, off_full, err_full = em.run(L, "full")
    s_hyb, off_hyb, err_hyb = em.run(L, "hybrid")
    assert off_full <= em.TOL and off_hyb <= em.TOL
    assert s_hyb <= s_full + 1, (s_hyb, s_full)
    assert err_hyb <= 5e-6 and err_full <= 5e-6
