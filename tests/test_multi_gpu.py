"""Multi-GPU parity (needs >= 2 GPUs; `gpurun --gpus 2 -- python -m pytest tests -m gpu`): one process per GPU under
torchrun, pixel rows band-sharded, the D / Gram / Phi^T y reductions over the library's NCCL communicator
(SURVEY 8e).  The assembled result must match the oracle exactly as the single-GPU path does."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _launch(tmp_path, world, W, H, ch, p_req, sampling, gs=0, affinity="bilateral"):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(tmp_path), str(W), str(H), str(ch), str(p_req),
           sampling, str(gs), affinity]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-4000:]
    return [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]


@pytest.mark.parametrize("world,W,H,ch,p_req,sampling,gs", [
    (2, 301, 203, 1, 120, "random", 0),            # ragged band boundary (203 rows over 2 ranks)
    (2, 256, 130, 3, 90, "spatially_uniform", 0),  # colour
    (2, 301, 203, 1, 120, "random", 1),            # with the orthonormalisation stage (Gram allreduce)
    (4, 320, 241, 1, 150, "random", 0),
    (8, 320, 241, 1, 150, "random", 0),
])
def test_band_sharded_matches_oracle(tmp_path, world, W, H, ch, p_req, sampling, gs):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import oracle_c as oc
    from oracle import oracle_np as o
    parts = _launch(tmp_path, world, W, H, ch, p_req, sampling, gs)
    img = o.synthetic_image(W, H, ch, seed=77)
    s = oc.random_sampling(W, H, p_req, 3) if sampling == "random" else oc.uniform_sampling(W, H, p_req)
    ref = o.run_pipeline(img, s, orthonormalise=bool(gs)) if gs else oc.run_pipeline(img, s)
    z = np.concatenate([p["z"] for p in parts], axis=0).astype(np.float64)
    assert z.shape == img.shape
    bands = [tuple(p["band"]) for p in parts]
    assert bands[0][0] == 0 and bands[-1][1] == H and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    for p in parts:
        assert np.array_equal(p["s"], s)                                   # sampling replicated, bit-exact
        assert np.array_equal(p["mu"], parts[0]["mu"])                     # replicated eigensolve: identical bits
        assert float(p["outside"][0]) == 0.0                               # a rank writes only its own band
    err_mu = float(np.max(np.abs(parts[0]["mu"] - ref["mu"]) / ref["mu"]))
    refz = np.asarray(ref["z"], dtype=np.float64)
    err_z = float(np.linalg.norm(z - refz) / np.linalg.norm(refz))
    err_dz = float(np.linalg.norm((z - img) - (refz - img)) / np.linalg.norm(refz - img))
    print(f"world={world} {W}x{H}x{ch} gs={gs}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= 1e-4 and err_z <= 1e-3 and err_dz <= 5e-3


def test_band_sharded_nlm_from_a_fresh_host_image(tmp_path):
    """gl_run with the NLM patch affinity on 2 GPUs: the 7x7 patches around band pixels and samples reach outside a rank's band,
    so every rank must have the whole image (the band-only upload left them uninitialised)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle_c as oc
    from oracle import oracle_np as o
    W, H, p_req = 120, 90, 60
    parts = _launch(tmp_path, 2, W, H, 1, p_req, "random", 0, "NLM")
    img = (o.synthetic_image(W, H, 1, seed=77) // 8 + 100).astype(np.uint8)
    s = oc.random_sampling(W, H, p_req, 3)
    ref = o.run_pipeline(img, s, kind=o.NLM, h_val=3.0)
    z = np.concatenate([p["z"] for p in parts], axis=0).astype(np.float64)
    assert np.array_equal(parts[0]["mu"], parts[1]["mu"])
    err_mu = float(np.max(np.abs(parts[0]["mu"] - ref["mu"]) / ref["mu"]))
    refz = np.asarray(ref["z"], dtype=np.float64)
    err_z = float(np.linalg.norm(z - refz) / np.linalg.norm(refz))
    err_dz = float(np.linalg.norm((z - img) - (refz - img)) / np.linalg.norm(refz - img))
    print(f"NLM world=2: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= 1e-4 and err_z <= 1e-3 and err_dz <= 5e-3


@pytest.mark.parametrize("tag,world", [("lion_rgb_photometric500", 2), ("barbara_uniform256", 2), ("cat_small_random50", 4)])
def test_golden_configs_band_sharded(tmp_path, tag, world):
    """BASELINE.json config 3 (input/lion.png as RGB, photometric affinity, p = 500) on 2 GPUs -- and configs 2 and 1 sharded as
    well -- against the same committed fixtures the single-GPU suite uses (tests/golden/make_golden.py)."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    parts = _launch(tmp_path, world, "golden:" + tag, 0, 0, 0, "-")
    g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    img = g["image"]
    src = np.repeat(img[:, :, None], 3, axis=2) if int(g["rgb"]) else img
    z = np.concatenate([p["z"] for p in parts], axis=0).astype(np.float64)
    assert z.shape == src.shape
    for p in parts:
        assert np.array_equal(p["s"], g["sample_indices"])
        assert np.array_equal(p["mu"], parts[0]["mu"])
        assert float(p["outside"][0]) == 0.0
    refz = g["z"].astype(np.float64)
    err_mu = float(np.max(np.abs(parts[0]["mu"] - g["mu"]) / g["mu"]))
    err_z = float(np.linalg.norm(z - refz) / np.linalg.norm(refz))
    err_dz = float(np.linalg.norm((z - src) - (refz - src)) / np.linalg.norm(refz - src))
    print(f"{tag} world={world}: err_mu={err_mu:.2e} err_z={err_z:.2e} err_dz={err_dz:.2e}")
    assert err_mu <= 1e-4 and err_z <= 1e-3 and err_dz <= 5e-3
