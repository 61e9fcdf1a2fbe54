"""The C host (image-processing-graph-laplacian_b200/hpc): the zlib-only PNG codec that replaces the reference's
read_png / write_png (hpc/read_img.c, hpc/write_img.c), checked against PIL on CPU, and -- on the GPU box -- the
image_processing binary itself (the drop-in for hpc/image_processing.c) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HPC = os.path.join(ROOT, "image-processing-graph-laplacian_b200", "hpc")
BIN = os.path.join(HPC, "image_processing")
PIL = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def png():
    path = os.path.join(HPC, "libglpng.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", HPC, "libglpng.so"])
    L = C.CDLL(path)
    rows_t = C.POINTER(C.POINTER(C.c_ubyte))
    L.read_png.argtypes = [C.c_char_p, C.POINTER(rows_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.read_png_rgb.argtypes = [C.c_char_p, C.POINTER(rows_t), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.write_png.argtypes = [C.c_char_p, rows_t, C.c_uint, C.c_uint]
    L.write_png_rgb.argtypes = [C.c_char_p, rows_t, C.c_uint, C.c_uint]
    return L


def _read(L, path, rgb=False):
    rows = C.POINTER(C.POINTER(C.c_ubyte))()
    w, h, col = C.c_int(), C.c_int(), C.c_int()
    rc = (L.read_png_rgb(path.encode(), C.byref(rows), C.byref(w), C.byref(h), C.byref(col)) if rgb
          else L.read_png(path.encode(), C.byref(rows), C.byref(w), C.byref(h)))
    if rc != 0:
        return rc, None
    nb = w.value * (3 if rgb else 1)
    out = np.stack([np.ctypeslib.as_array(rows[i], shape=(nb,)).copy() for i in range(h.value)])
    return 0, out.reshape(h.value, w.value, 3) if rgb else out


def _write(L, path, a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    h = a.shape[0]
    ptrs = (C.POINTER(C.c_ubyte) * h)(*[a[i].ctypes.data_as(C.POINTER(C.c_ubyte)) for i in range(h)])
    fn = L.write_png_rgb if a.ndim == 3 else L.write_png
    return fn(path.encode(), ptrs, a.shape[1], h)


def _libpng_gray(rgb):
    """png_set_rgb_to_gray(png, 1, -1, -1) as the reference calls it (hpc/read_img.c:49): libpng's default integer
    weights over 32768, truncated; grey pixels pass through."""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    v = (6968 * r + 23434 * g + 2366 * b) >> 15
    same = (r == g) & (g == b)
    return np.where(same, r, v).astype(np.uint8)


def test_png_grey_roundtrip_and_pil(png, tmp_path):
    rng = np.random.RandomState(0)
    for (h, w) in ((1, 1), (7, 13), (100, 100), (233, 350)):
        a = rng.randint(0, 256, size=(h, w)).astype(np.uint8)
        a[: h // 2] = (np.arange(w) * 3 % 256).astype(np.uint8)          # smooth part: exercises the Up filter
        ours, pil = str(tmp_path / "ours.png"), str(tmp_path / "pil.png")
        assert _write(png, ours, a) == 0
        assert np.array_equal(np.asarray(PIL.open(ours)), a)              # PIL reads what write_png wrote
        PIL.fromarray(a).save(pil)                                        # PIL picks adaptive filters (all five types)
        rc, b = _read(png, pil)
        assert rc == 0 and np.array_equal(b, a)


def test_png_colour_types(png, tmp_path):
    rng = np.random.RandomState(1)
    rgb = rng.randint(0, 256, size=(31, 45, 3)).astype(np.uint8)
    rgb[5:9, :, 1] = rgb[5:9, :, 0]
    rgb[5:9, :, 2] = rgb[5:9, :, 0]                                       # a grey band inside a colour file
    alpha = rng.randint(0, 256, size=(31, 45, 1)).astype(np.uint8)
    p = str(tmp_path / "x.png")
    PIL.fromarray(rgb).save(p)
    rc, g = _read(png, p)
    assert rc == 0 and np.array_equal(g, _libpng_gray(rgb))
    rc, c = _read(png, p, rgb=True)
    assert rc == 0 and np.array_equal(c, rgb)
    PIL.fromarray(np.concatenate([rgb, alpha], axis=2), "RGBA").save(p)   # alpha dropped
    rc, g = _read(png, p)
    assert rc == 0 and np.array_equal(g, _libpng_gray(rgb))
    grey = rgb[..., 0]
    PIL.fromarray(np.concatenate([grey[..., None], alpha], axis=2), "LA").save(p)   # grey+alpha (bear.png's type)
    rc, g = _read(png, p)
    assert rc == 0 and np.array_equal(g, grey)
    PIL.fromarray(rgb).quantize(16).save(p)                               # palette, 4 bits per index
    pal = np.asarray(PIL.open(p).convert("RGB"))
    rc, c = _read(png, p, rgb=True)
    assert rc == 0 and np.array_equal(c, pal)
    PIL.fromarray((grey.astype(np.uint16) << 8) | 7).save(p)              # 16-bit grey keeps the high byte
    rc, g = _read(png, p)
    assert rc == 0 and np.array_equal(g, grey)
    PIL.fromarray(grey > 127).save(p)                                     # 1-bit grey scales to 0/255
    rc, g = _read(png, p)
    assert rc == 0 and np.array_equal(g, np.where(grey > 127, 255, 0).astype(np.uint8))
    assert _write(png, p, rgb) == 0 and np.array_equal(np.asarray(PIL.open(p)), rgb)


def test_png_errors(png, tmp_path, capfd):
    rc, _ = _read(png, str(tmp_path / "missing.png"))
    assert rc == -1
    assert "Could not open file" in capfd.readouterr().err               # hpc/read_img.c:16
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    assert _read(png, str(bad))[0] == -1
    good = str(tmp_path / "good.png")
    PIL.fromarray(np.zeros((9, 9), np.uint8)).save(good)
    data = bytearray(open(good, "rb").read())
    data[-20] ^= 0xFF                                                     # corrupt the IDAT: CRC must catch it
    bad.write_bytes(bytes(data))
    assert _read(png, str(bad))[0] == -1
    assert _write(png, str(tmp_path / "no_such_dir" / "x.png"), np.zeros((2, 2), np.uint8)) == -1


def test_reference_inputs_decode_like_pil(png):
    """The ten images the reference ships (only present in the build container)."""
    d = "/root/reference/input"
    if not os.path.isdir(d):
        pytest.skip("reference inputs not on this machine")
    for f in sorted(os.listdir(d)):
        im = PIL.open(os.path.join(d, f))
        rc, g = _read(png, os.path.join(d, f))
        assert rc == 0, f
        if im.mode == "L":
            assert np.array_equal(g, np.asarray(im)), f
        elif im.mode == "LA":
            assert np.array_equal(g, np.asarray(im)[..., 0]), f
        else:
            assert np.array_equal(g, _libpng_gray(np.asarray(im.convert("RGB")))), f


def test_host_binary_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", HPC])
    p = str(tmp_path / "in.png")
    PIL.fromarray(np.zeros((16, 16), np.uint8)).save(p)
    r = subprocess.run([BIN, "-f", p], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_shared_buffer_handover_between_ranks(tmp_path):
    """hpc/glshare.h: the protocol by which the forked ranks fill the shared output image -- and, one after the other,
    every eigenvector column dump -- and rank 0 reads it.  Plain processes, no GPU: 4 ranks, 400 hand-overs with random
    delays; rank 0 must see exactly the current generation in every slot."""
    exe = str(tmp_path / "share_harness")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", HPC, os.path.join(os.path.dirname(os.path.abspath(__file__)), "share_harness.c"),
                           "-o", exe])
    for size in (1, 2, 4):
        r = subprocess.run([exe, str(size), "400"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.strip() == "ok", (size, r.stdout, r.stderr)


# ---------------------------------------------------------------------------------------------------------------
# the binary on a GPU
# ---------------------------------------------------------------------------------------------------------------
def _run_bin(tmp_path, args):
    r = subprocess.run([BIN] + args, capture_output=True, text=True, cwd=str(tmp_path), timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["test_uniform100", "cat_small_uniform_m20"])
def test_binary_matches_oracle(golden, tmp_path, tag):
    from oracle import oracle_np as o
    g = golden(tag)
    img = g["image"]
    src = str(tmp_path / "in.png")
    PIL.fromarray(img).save(src)
    args = ["-f", src, "-sample_size", str(int(g["p_req"])), "-num_eigvals", str(int(g["m"]))]
    r = _run_bin(tmp_path, args)
    out = r.stdout
    # the reference's stdout vocabulary (hpc/image_processing.c:193-269,285,296,314)
    for line in ("Running with 1 processes", "Read image", "Sample size: %d" % len(g["sample_indices"]),
                 "Computing affinity matrices... ", "Computing Laplacian matrices... ",
                 "Computing %d smallest eigenvalues... (epsilon: 0.1)" % int(g["m"]), "approximation... ",
                 "Computing output image... ", "Total computation time: "):
        assert line in out, (line, out)
    assert np.array_equal(np.asarray(PIL.open(str(tmp_path / "results" / "input.png"))), img)
    z8 = np.asarray(PIL.open(str(tmp_path / "results" / "output.png"))).astype(np.int32)
    ref = o.quantise(g["z"]).astype(np.int32)
    assert z8.shape == ref.shape
    assert np.max(np.abs(z8 - ref)) <= 1 and np.mean(z8 != ref) < 0.01    # truncation at integer boundaries only
    # eigenvalue dump (WriteDiagMat, hpc/image_processing.c:234): PETSc ASCII Vec layout, values in ascending order
    lines = open(str(tmp_path / "results" / "eigenvalues_laplacian.txt")).read().split("\n")
    assert lines[0].startswith("Vec Object:")
    mu = np.array([float(x) for x in lines[2:] if x.strip()])
    assert mu.shape == g["mu"].shape and np.max(np.abs(mu - g["mu"]) / g["mu"]) <= 1e-4


@pytest.mark.gpu
def test_binary_options(golden, tmp_path):
    """-use_slepc, -filter_pow / -filter_gain, -sampling random -seed, -synthetic, -color, -gram_schmidt, -o."""
    from oracle import oracle_c as oc
    from oracle import oracle_np as o
    g = golden("cat_small_random50")
    img = g["image"]
    src = str(tmp_path / "in.png")
    PIL.fromarray(img).save(src)
    out = str(tmp_path / "o.png")
    _run_bin(tmp_path, ["-f", src, "-sampling", "random", "-seed", str(int(g["seed"])), "-sample_size", "50", "-use_slepc", "-o", out])
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert np.max(np.abs(z8 - o.quantise(g["z"]).astype(np.int32))) <= 1
    _run_bin(tmp_path, ["-f", src, "-sampling", "random", "-seed", str(int(g["seed"])), "-sample_size", "50", "-gram_schmidt", "-o", out])
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert np.max(np.abs(z8 - o.quantise(g["z_gs"]).astype(np.int32))) <= 1
    s = g["sample_indices"]
    ref = o.run_pipeline(img, s, gain=-1.0, power=2.0)
    _run_bin(tmp_path, ["-f", src, "-sampling", "random", "-seed", str(int(g["seed"])), "-sample_size", "50",
                        "-filter_gain", "-1", "-filter_pow", "2", "-o", out])
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert np.max(np.abs(z8 - o.quantise(ref["z"]).astype(np.int32))) <= 1
    # synthetic colour image generated on the device, RGB affinity, RGB output
    r = _run_bin(tmp_path, ["-synthetic", "160x96", "-color", "-sample_size", "60", "-o", out])
    assert "Read image synthetic 160x96 of size 160x96" in r.stdout
    simg = o.synthetic_image(160, 96, 3, seed=1234)
    assert np.array_equal(np.asarray(PIL.open(str(tmp_path / "results" / "input.png"))), simg)
    ref = oc.run_pipeline(simg, oc.uniform_sampling(160, 96, 60))
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert z8.shape == (96, 160, 3) and np.max(np.abs(z8 - o.quantise(ref["z"]).astype(np.int32))) <= 1
    # missing -f: message and exit(1) (hpc/image_processing.c:88-92)
    r = subprocess.run([BIN], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 1 and "No filename found (option -f)" in r.stderr
    # -no_approx (hpc/image_processing.c:155-181): the full-matrix mode, matrix-free here; small image, narrow kernel
    small = o.synthetic_image(72, 56, 1, seed=9)
    PIL.fromarray(small).save(src)
    r = _run_bin(tmp_path, ["-f", src, "-no_approx", "-h_loc", "5", "-o", out])
    for line in ("Computing entire affinity matrix... ", "Computing entire Laplacian matrix... ", "Computing output image... "):
        assert line in r.stdout
    ref = o.run_full(small, "bilateral", 5.0, 30.0)
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert np.max(np.abs(z8 - ref["z"].astype(np.uint8).astype(np.int32))) <= 1


def _read_vec(path):
    lines = open(path).read().split("\n")
    assert lines[0].startswith("Vec Object:")
    return np.array([float(x) for x in lines[2:] if x.strip()])


def _same_up_to_sign(a, b, tol):
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) <= tol * np.linalg.norm(b)


@pytest.mark.gpu
def test_binary_eigenvector_dumps(golden, tmp_path):
    """-dump_eigvecs K: WriteMatCol / WritePngMatCol of the extrapolated eigenvectors (hpc/image_processing.c:255-260,
    hpc/display.c:85-126): text in the PETSc ASCII Vec layout, PNGs with the reference's byte cast or (-dump_scaled)
    stretched to the column's range.  Columns are compared with the oracle's Phi up to sign."""
    from oracle import oracle_np as o
    g = golden("test_uniform100")
    img = g["image"]
    src = str(tmp_path / "in.png")
    PIL.fromarray(img).save(src)
    ref = o.run_pipeline(img, g["sample_indices"], return_phi=True)
    _run_bin(tmp_path, ["-f", src, "-sample_size", str(int(g["p_req"])), "-dump_eigvecs", "3", "-dump_scaled"])
    for k in range(3):
        v = _read_vec(str(tmp_path / "results" / ("eigenvector_%d_laplacian.txt" % k)))
        assert v.shape == (img.size,)
        assert _same_up_to_sign(v, ref["phi"][:, k], 2e-3), k              # fp16 storage of Phi
        png = np.asarray(PIL.open(str(tmp_path / "results" / ("eigenvector_%d_laplacian.png" % k)))).astype(np.float64)
        assert png.shape == img.shape and png.min() == 0 and png.max() == 255
        want = np.floor((v - v.min()) * (255.0 / (v.max() - v.min()))).reshape(img.shape)
        assert np.max(np.abs(png - want)) <= 1
    # without -dump_scaled: the reference's cast of O(1/sqrt(n)) entries gives a black image
    _run_bin(tmp_path, ["-f", src, "-sample_size", str(int(g["p_req"])), "-dump_eigvecs", "1"])
    assert np.asarray(PIL.open(str(tmp_path / "results" / "eigenvector_0_laplacian.png"))).max() <= 1
    # the output image is the same with and without the dumps (Phi materialised early for them)
    z8 = np.asarray(PIL.open(str(tmp_path / "results" / "output.png"))).astype(np.int32)
    assert np.max(np.abs(z8 - o.quantise(g["z"]).astype(np.int32))) <= 1


@pytest.mark.gpu
def test_binary_two_ranks(golden, tmp_path):
    """-ngpus 2: one forked process per GPU (the reference's `mpiexec -n 2`), bands of rows, NCCL id through shared
    memory, both ranks writing their band into the shared output image."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle_np as o
    g = golden("barbara_uniform256")
    src, out = str(tmp_path / "in.png"), str(tmp_path / "o.png")
    PIL.fromarray(g["image"]).save(src)
    r = _run_bin(tmp_path, ["-f", src, "-sample_size", "256", "-ngpus", "2", "-o", out])
    assert "Running with 2 processes" in r.stdout
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert np.max(np.abs(z8 - o.quantise(g["z"]).astype(np.int32))) <= 1
    r = _run_bin(tmp_path, ["-f", src, "-sample_size", "256", "-ngpus", "2", "-gram_schmidt", "-o", out])
    z8 = np.asarray(PIL.open(out)).astype(np.int32)
    assert np.max(np.abs(z8 - o.quantise(g["z_gs"]).astype(np.int32))) <= 1
    # eigenvector dump gathered from both bands equals the one-rank dump
    _run_bin(tmp_path, ["-f", src, "-sample_size", "256", "-ngpus", "2", "-dump_eigvecs", "2", "-o", out])
    two = [_read_vec(str(tmp_path / "results" / ("eigenvector_%d_laplacian.txt" % k))) for k in range(2)]
    _run_bin(tmp_path, ["-f", src, "-sample_size", "256", "-dump_eigvecs", "2", "-o", out])
    for k in range(2):
        one = _read_vec(str(tmp_path / "results" / ("eigenvector_%d_laplacian.txt" % k)))
        assert one.shape == (g["image"].size,) and _same_up_to_sign(two[k], one, 1e-3)
