"""Python host-side mirror of the reference's Python interface
(python/image_processing.py:244 `image_processing(y, cr, cb, **kwargs)` with the plugin dictionaries
python/sampling/__init__.py:6-9 and python/affinity_methods/__init__.py:8-13), running on libglcuda.so
through its C ABI (include/gl_cuda.h) with ctypes.  No numpy/torch compute on the path: numpy only
carries host buffers.  There is NO CPU fallback -- creating a Context without a B200 raises.

Import as `ipgl_b200` (repo-root shim) because this directory's name is not a Python identifier.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglcuda.so")

# gl_status
OK, ERR_CUDA, ERR_ARG, ERR_NOMEM, ERR_NCCL, ERR_UNSUPPORTED, ERR_NOTCONVERGED = range(7)
# gl_affinity_kind -- names follow python/affinity_methods/__init__.py:6
SPATIAL, PHOTOMETRIC, BILATERAL, NLM = "spatial", "photometric", "bilateral", "NLM"     # python/affinity_methods/__init__.py:6
AFFINITY_KINDS = {BILATERAL: 0, PHOTOMETRIC: 1, SPATIAL: 2, NLM: 3, "nlm": 3}
NLM_H = 3.0    # python/affinity_methods/NLM.py:12
# sampling -- names follow python/sampling/__init__.py:4
RANDOM, SPATIALLY_UNIFORM = "random", "spatially_uniform"
# gl_mat_kind
MAT_KA, MAT_KB, MAT_EIGVEC, MAT_DIAG, MAT_PHI, MAT_FULL = 1, 2, 3, 4, 5, 6
STAGES = ["h2d", "sampling", "affinity", "laplacian", "eigen", "nystroem", "gram_schmidt", "filter", "d2h", "total",
          "k_affinity_b", "k_gemm", "k_filter_project", "k_filter_apply", "k_jacobi"]

EXPORTS = [
    "gl_version", "gl_last_error", "gl_default_params", "gl_device_count", "gl_memory_stats", "gl_kernel_launches",
    "gl_ctx_create", "gl_ctx_destroy", "gl_ctx_sync", "gl_ctx_stage_ms", "gl_ctx_set_option", "gl_ctx_mark",
    "gl_ctx_mark_elapsed_ms", "gl_ctx_flush_l2",
    "gl_comm_unique_id", "gl_comm_init",
    "gl_set_image", "gl_set_image_rows", "gl_set_synthetic_image", "gl_get_image", "gl_get_band",
    "gl_sampling_uniform", "gl_sampling_random", "gl_set_samples", "gl_get_samples",
    "gl_affinity", "gl_laplacian", "gl_eigensolve", "gl_inverse_iteration", "gl_nystroem", "gl_nystroem_filter", "gl_orthonormalise", "gl_filter",
    "gl_diag_inverse", "gl_diag_pow", "gl_full_affinity", "gl_full_laplacian", "gl_full_result", "gl_run", "gl_run_resident",
    "gl_mat_info_get", "gl_mat_retain", "gl_mat_destroy", "gl_mat_download", "gl_mat_download_cols", "gl_mat_rowsums", "gl_mat_upload",
    "gl_host_alloc", "gl_host_free", "gl_kb_layout_host",
    "gl_sinkhorn", "gl_orthogonalisation", "gl_smoothing_matrix", "gl_matrix_filter",
]


class GLError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libglcuda status {status}: {msg}")
        self.status = status


class Params(C.Structure):
    _fields_ = [("affinity_kind", C.c_int), ("h_loc", C.c_double), ("h_val", C.c_double),
                ("sampling_random", C.c_int), ("seed", C.c_uint32), ("sample_size", C.c_uint),
                ("num_eigvals", C.c_int), ("gain", C.c_double), ("power", C.c_double),
                ("gram_schmidt", C.c_int), ("clip_low", C.c_int)]


class MatInfo(C.Structure):
    _fields_ = [("kind", C.c_int), ("rows", C.c_int64), ("cols", C.c_int64), ("local_rows", C.c_int64),
                ("ld", C.c_int64), ("elem_bytes", C.c_int), ("scale", C.c_double), ("stored_blocks", C.c_int64),
                ("layout", C.c_int), ("stored_pairs", C.c_int64), ("mma_pairs", C.c_int64)]


_lib = None


def build(verbose=False):
    """Compile libglcuda.so in-tree (nvcc, sm_100a)."""
    cmd = ["make", "-C", _HERE, "-j", str(os.cpu_count() or 4)]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GLError(ERR_CUDA, f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.gl_last_error.restype = C.c_char_p
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.gl_ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
        L.gl_ctx_destroy.argtypes = [vp]
        L.gl_ctx_sync.argtypes = [vp]
        L.gl_ctx_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.gl_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.gl_kernel_launches.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.gl_memory_stats.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int]
        L.gl_ctx_mark.argtypes = [vp, C.c_int]
        L.gl_ctx_mark_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.gl_ctx_flush_l2.argtypes = [vp, C.c_size_t]
        L.gl_comm_unique_id.argtypes = [vp]
        L.gl_comm_init.argtypes = [vp, vp]
        L.gl_set_image.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.gl_set_synthetic_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.gl_get_image.argtypes = [vp, vp]
        L.gl_get_band.argtypes = [vp, ip, ip]
        L.gl_sampling_uniform.argtypes = [vp, C.c_uint, C.POINTER(C.c_uint)]
        L.gl_sampling_random.argtypes = [vp, C.c_uint, C.c_uint32, C.POINTER(C.c_uint)]
        L.gl_set_samples.argtypes = [vp, vp, C.c_uint]
        L.gl_get_samples.argtypes = [vp, vp, C.c_uint, C.POINTER(C.c_uint)]
        L.gl_affinity.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.POINTER(vp), C.POINTER(vp)]
        L.gl_laplacian.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]
        L.gl_eigensolve.argtypes = [vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.gl_inverse_iteration.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), ip,
                                           C.POINTER(C.c_double)]
        L.gl_nystroem.argtypes = [vp, vp, vp, vp, C.POINTER(vp)]
        L.gl_nystroem_filter.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_int, C.POINTER(vp), vp, vp]
        L.gl_orthonormalise.argtypes = [vp, vp, vp]
        L.gl_filter.argtypes = [vp, vp, vp, C.c_double, C.c_int, vp, vp]
        L.gl_sinkhorn.argtypes = [vp, vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
        L.gl_orthogonalisation.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]
        L.gl_smoothing_matrix.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]
        L.gl_matrix_filter.argtypes = [vp, vp, vp, vp, C.c_int, vp]
        L.gl_full_affinity.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
        L.gl_full_laplacian.argtypes = [vp, vp, C.POINTER(vp)]
        L.gl_full_result.argtypes = [vp, vp, vp, vp]
        L.gl_diag_inverse.argtypes = [vp, vp, C.POINTER(vp)]
        L.gl_diag_pow.argtypes = [vp, vp, C.c_double, C.POINTER(vp)]
        L.gl_run.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(Params), vp, vp, C.POINTER(C.c_uint), ip, vp, C.c_size_t]
        L.gl_run_resident.argtypes = [vp, C.POINTER(Params), vp, vp, C.POINTER(C.c_uint), ip, vp, C.c_size_t]
        L.gl_mat_info_get.argtypes = [vp, C.POINTER(MatInfo)]
        L.gl_mat_destroy.argtypes = [vp]
        L.gl_mat_retain.argtypes = [vp]
        L.gl_mat_download.argtypes = [vp, vp, vp, C.c_size_t]
        L.gl_mat_download_cols.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_size_t]
        L.gl_mat_rowsums.argtypes = [vp, vp, vp, C.c_size_t]
        L.gl_mat_upload.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int64, C.POINTER(vp)]
        L.gl_kb_layout_host.argtypes = [C.c_int, C.c_int64, C.c_int64, vp, C.c_uint, C.c_double, C.c_int, C.c_int, C.c_int, ip,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, vp, vp, vp]
        L.gl_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.gl_host_free.argtypes = [vp]
        L.gl_default_params.argtypes = [C.POINTER(Params)]
        L.gl_device_count.argtypes = [ip]
        _lib = L
    return _lib


def _check(status):
    if status != OK:
        raise GLError(status, lib().gl_last_error().decode(errors="replace"))


def default_params(**kw) -> Params:
    p = Params()
    lib().gl_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "affinity":
            p.affinity_kind = AFFINITY_KINDS[v]
            if p.affinity_kind == AFFINITY_KINDS[NLM] and "h_val" not in kw:
                p.h_val = NLM_H            # the patch kernel's h (NLM.py:12); h_loc is unused
        elif k == "sampling":
            p.sampling_random = 1 if v == RANDOM else 0
        else:
            if not hasattr(p, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(p, k, v)
    return p


def kb_layout(width, q0, q1, samples, h_loc=40.0, cutoff=True, strips=0, block=64):
    """Host-only: the block layout gl_affinity gives K_B (gl_kb_layout_host).  Returns a dict with strips, tile_first,
    tile_count, starts, perm."""
    s = np.ascontiguousarray(samples, dtype=np.uint32)
    S, nt, nb = C.c_int(), C.c_int64(), C.c_int64()
    _check(lib().gl_kb_layout_host(width, q0, q1, s.ctypes.data, len(s), h_loc, int(cutoff), strips, block, C.byref(S), C.byref(nt),
                                   C.byref(nb), None, None, None, None))
    first, count = np.empty(nt.value, np.int32), np.empty(nt.value, np.int32)
    starts = np.empty(nb.value, np.int32)
    perm = np.empty((len(s) + 63) // 64 * 64 + 64, np.uint32)
    _check(lib().gl_kb_layout_host(width, q0, q1, s.ctypes.data, len(s), h_loc, int(cutoff), strips, block, C.byref(S), C.byref(nt),
                                   C.byref(nb), first.ctypes.data, count.ctypes.data, starts.ctypes.data, perm.ctypes.data))
    return dict(strips=S.value, tile_first=first, tile_count=count, starts=starts, perm=perm, n_blocks=nb.value, block=block)


def device_count() -> int:
    n = C.c_int()
    _check(lib().gl_device_count(C.byref(n)))
    return n.value


class PinnedArray:
    """numpy view over page-locked host memory (for end-to-end timing with async copies)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = C.c_void_p()
        _check(lib().gl_host_alloc(C.byref(self._ptr), max(self.nbytes, 1)))
        buf = (C.c_char * self.nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().gl_host_free(self._ptr)
            self._ptr = None


class Mat:
    """Opaque device matrix (the reference's Mat/Vec)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        ctx._mats.add(self)

    @property
    def info(self) -> MatInfo:
        i = MatInfo()
        _check(lib().gl_mat_info_get(self.h, C.byref(i)))
        return i

    def download(self) -> np.ndarray:
        i = self.info
        if i.kind == MAT_KB:
            shape = (i.local_rows, i.rows)          # stored transposed: band pixels x p
        elif i.kind == MAT_PHI:
            shape = (i.local_rows, i.cols)
        elif i.kind == MAT_DIAG:
            shape = (i.rows,)
        else:
            shape = (i.rows, i.cols)
        out = np.empty(shape, dtype=np.float64)
        _check(lib().gl_mat_download(self.ctx.h, self.h, out.ctypes.data, out.size))
        return out

    def download_cols(self, col0: int, ncols: int = 1) -> np.ndarray:
        """columns [col0, col0 + ncols) of a Phi (band rows), eigenvector or p x p matrix (WriteMatCol, hpc/display.c:85-100)"""
        i = self.info
        rows = i.local_rows if i.kind == MAT_PHI else i.rows
        out = np.empty((rows, ncols), dtype=np.float64)
        _check(lib().gl_mat_download_cols(self.ctx.h, self.h, col0, ncols, out.ctypes.data, out.size))
        return out

    def rowsums(self) -> np.ndarray:
        out = np.empty(self.info.rows, dtype=np.float64)
        _check(lib().gl_mat_rowsums(self.ctx.h, self.h, out.ctypes.data, out.size))
        return out

    def destroy(self):
        # a handle must not outlive its context (the buffers go back to the context's allocator)
        if self.h and self.ctx is not None and self.ctx.h:
            lib().gl_mat_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Context:
    """One GPU (one rank).  world > 1: create one Context per process/GPU and call init_comm()."""

    def __init__(self, device=0, rank=0, world=1):
        self.h = C.c_void_p()
        _check(lib().gl_ctx_create(C.byref(self.h), device, rank, world))
        self.rank, self.world = rank, world
        self.shape = None
        self._mats = weakref.WeakSet()

    def close(self):
        if self.h:
            for m in list(self._mats):
                m.destroy()
            lib().gl_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key, value):
        _check(lib().gl_ctx_set_option(self.h, key.encode(), str(value).encode()))

    def sync(self):
        _check(lib().gl_ctx_sync(self.h))

    def stage_ms(self) -> dict:
        a = (C.c_float * len(STAGES))()
        _check(lib().gl_ctx_stage_ms(self.h, a))
        return dict(zip(STAGES, list(a)))

    def mark(self, slot):
        _check(lib().gl_ctx_mark(self.h, slot))

    def flush_l2(self, nbytes=0):
        """write scratch (default: twice the L2 size) on the context stream: evicts the L2 between timed iterations"""
        _check(lib().gl_ctx_flush_l2(self.h, nbytes))

    def elapsed_ms(self, a, b) -> float:
        ms = C.c_float()
        _check(lib().gl_ctx_mark_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def memory_stats(self, reset_peak=False) -> dict:
        """bytes of device memory held by live handles / cached for reuse / peak live since the last reset"""
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(lib().gl_memory_stats(self.h, C.byref(a), C.byref(b), C.byref(c), int(reset_peak)))
        return dict(live=a.value, cached=b.value, peak=c.value)

    def kernel_launches(self) -> int:
        n = C.c_longlong()
        _check(lib().gl_kernel_launches(self.h, C.byref(n)))
        return n.value

    # ---- NCCL bootstrap -------------------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().gl_comm_unique_id(buf))
        return buf.raw

    def init_comm(self, uid: bytes):
        _check(lib().gl_comm_init(self.h, C.create_string_buffer(uid, 128)))

    # ---- image / samples ------------------------------------------------------------------
    def set_image(self, img: np.ndarray):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        H, W = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        _check(lib().gl_set_image(self.h, img.ctypes.data, W, H, ch))
        self.sync()  # `img` may be a temporary
        self.shape = (H, W, ch)

    def set_synthetic_image(self, width, height, channels=1, seed=1234):
        _check(lib().gl_set_synthetic_image(self.h, width, height, channels, seed))
        self.shape = (height, width, channels)

    def get_image(self) -> np.ndarray:
        H, W, ch = self.shape
        out = np.empty((H, W, ch), dtype=np.uint8)
        _check(lib().gl_get_image(self.h, out.ctypes.data))
        return out[:, :, 0] if ch == 1 else out

    def band(self):
        a, b = C.c_int(), C.c_int()
        _check(lib().gl_get_band(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sampling(self, method, requested, seed=0) -> np.ndarray:
        n = C.c_uint()
        if method == RANDOM:
            _check(lib().gl_sampling_random(self.h, requested, seed, C.byref(n)))
        else:
            _check(lib().gl_sampling_uniform(self.h, requested, C.byref(n)))
        return self.get_samples()

    def set_samples(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        _check(lib().gl_set_samples(self.h, idx.ctypes.data, len(idx)))

    def get_samples(self) -> np.ndarray:
        n = C.c_uint()
        _check(lib().gl_get_samples(self.h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint32)
        _check(lib().gl_get_samples(self.h, out.ctypes.data, n.value, C.byref(n)))
        return out

    # ---- stages (names follow the reference's hpc/ entry points) -----------------------------
    def affinity(self, kind=BILATERAL, h_loc=40.0, h_val=None):
        """h_val: the photometric bandwidth (30, hpc/affinity.c:117) or, for NLM, the patch kernel's h (3, NLM.py:12)"""
        if h_val is None:
            h_val = NLM_H if AFFINITY_KINDS[kind] == AFFINITY_KINDS[NLM] else 30.0
        a, b = C.c_void_p(), C.c_void_p()
        _check(lib().gl_affinity(self.h, AFFINITY_KINDS[kind], h_loc, h_val, C.byref(a), C.byref(b)))
        return Mat(self, a), Mat(self, b)

    def laplacian(self, K_A: Mat, K_B: Mat):
        a, b = C.c_void_p(), C.c_void_p()
        _check(lib().gl_laplacian(self.h, K_A.h, K_B.h, C.byref(a), C.byref(b)))
        return Mat(self, a), Mat(self, b)

    def eigensolve(self, L_A: Mat, m=-1):
        u, d, di = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().gl_eigensolve(self.h, L_A.h, m, C.byref(u), C.byref(d), C.byref(di)))
        return Mat(self, u), Mat(self, d), Mat(self, di)

    def inverse_iteration(self, L_A: Mat, m=-1, opti_gs=1, epsilon=0.1, max_iterations=1000):
        """InversePowerIteration (hpc/inverse_power_it.c:86-252): returns (eigvecs, eigvals, eigvals_inv, iterations, residual)."""
        u, d, di = C.c_void_p(), C.c_void_p(), C.c_void_p()
        it, res = C.c_int(), C.c_double()
        _check(lib().gl_inverse_iteration(self.h, L_A.h, m, opti_gs, epsilon, max_iterations, C.byref(u), C.byref(d), C.byref(di),
                                          C.byref(it), C.byref(res)))
        return Mat(self, u), Mat(self, d), Mat(self, di), it.value, res.value

    def nystroem(self, L_B: Mat, phi_A: Mat, eigvals_inv: Mat):
        p = C.c_void_p()
        _check(lib().gl_nystroem(self.h, L_B.h, phi_A.h, eigvals_inv.h, C.byref(p)))
        return Mat(self, p)

    def nystroem_filter(self, L_B: Mat, phi_A: Mat, eigvals_inv: Mat, f_eigvals: Mat, gain=3.0, clip_low=False, keep_phi=True):
        """Nystroem + ComputeResultFromLaplacian in one pass over Phi; returns (phi, z) -- phi is None with keep_phi=False
        (Phi is then never written to memory)."""
        H, W, ch = self.shape
        z = np.zeros((H, W, ch), dtype=np.float32)
        p = C.c_void_p()
        _check(lib().gl_nystroem_filter(self.h, L_B.h, phi_A.h, eigvals_inv.h, f_eigvals.h, gain, int(clip_low),
                                        C.byref(p) if keep_phi else None, z.ctypes.data, None))
        return (Mat(self, p) if keep_phi else None), (z[:, :, 0] if ch == 1 else z)

    def orthonormalise(self, phi: Mat) -> np.ndarray:
        norms = np.empty(phi.info.cols, dtype=np.float64)
        _check(lib().gl_orthonormalise(self.h, phi.h, norms.ctypes.data))
        return norms

    def diag_pow(self, d: Mat, power) -> Mat:
        o = C.c_void_p()
        _check(lib().gl_diag_pow(self.h, d.h, power, C.byref(o)))
        return Mat(self, o)

    def diag_inverse(self, d: Mat) -> Mat:
        o = C.c_void_p()
        _check(lib().gl_diag_inverse(self.h, d.h, C.byref(o)))
        return Mat(self, o)

    # ---- the prototype's experimental blocks (python/image_processing.py:90-241; csrc/proto.cu) ---------------
    def sinkhorn(self, phi: Mat, Pi: Mat, iterations=100):
        """sinkhorn(phi, Pi) (:90-107) on a stored Phi in raster order: returns the handles (W_A [p x p], W_ABt [band pixels x p]);
        W_ABt.download()[j, i] is the prototype's W_AB[i, j] with its columns back in raster order."""
        a, b = C.c_void_p(), C.c_void_p()
        _check(lib().gl_sinkhorn(self.h, phi.h, Pi.h, iterations, C.byref(a), C.byref(b)))
        return Mat(self, a), Mat(self, b)

    def orthogonalisation(self, K_A: Mat, K_B: Mat):
        """orthogonalisation(A, B) (:110-127): returns the handles (V [band pixels x p, raster order], Pi [p])."""
        v, d = C.c_void_p(), C.c_void_p()
        _check(lib().gl_orthogonalisation(self.h, K_A.h, K_B.h, C.byref(v), C.byref(d)))
        return Mat(self, v), Mat(self, d)

    def smoothing_matrix(self, phi: Mat, Pi: Mat):
        """smoothing_matrix(sample_indices, phi, Pi) (:151-194): returns the handles (V [band pixels x p, raster order], L [p])."""
        v, d = C.c_void_p(), C.c_void_p()
        _check(lib().gl_smoothing_matrix(self.h, phi.h, Pi.h, C.byref(v), C.byref(d)))
        return Mat(self, v), Mat(self, d)

    def matrix_filter(self, V: Mat, L: Mat, coef):
        """z = sum_k coef[k] (V diag(L) V^T)^k y for the current image, float32 [H, W(, C)], not clipped."""
        H, W, ch = self.shape
        z = np.zeros((H, W, ch), dtype=np.float32)
        c = np.ascontiguousarray(coef, dtype=np.float64)
        _check(lib().gl_matrix_filter(self.h, V.h, L.h, c.ctypes.data, len(c), z.ctypes.data))
        return z[:, :, 0] if ch == 1 else z

    def smoothing(self, phi: Mat, Pi: Mat):
        """smoothing(y, sample_indices, phi, Pi) (:197-219): z = V L V^T y with (V, L) = smoothing_matrix."""
        V, L = self.smoothing_matrix(phi, Pi)
        return self.matrix_filter(V, L, [0.0, 1.0])

    def sharpening(self, phi: Mat, Pi: Mat, beta=1.5):
        """sharpening(y, sample_indices, phi, Pi) (:222-241): z = (1 + beta) W^2 y - beta W^3 y, W = V L V^T, beta = 1.5 (:231)."""
        V, L = self.smoothing_matrix(phi, Pi)
        return self.matrix_filter(V, L, [0.0, 0.0, 1.0 + beta, -beta])

    # ---- the reference's -no_approx mode (matrix-free) -----------------------------------------------
    def full_affinity(self, kind=BILATERAL, h_loc=40.0, h_val=30.0) -> Mat:
        k = C.c_void_p()
        _check(lib().gl_full_affinity(self.h, AFFINITY_KINDS[kind], h_loc, h_val, C.byref(k)))
        return Mat(self, k)

    def full_laplacian(self, K: Mat) -> Mat:
        o = C.c_void_p()
        _check(lib().gl_full_laplacian(self.h, K.h, C.byref(o)))
        return Mat(self, o)

    def full_result(self, L: Mat, want_u8=False):
        H, W, ch = self.shape
        z = np.zeros((H, W, ch), dtype=np.float32)
        z8 = np.zeros((H, W, ch), dtype=np.uint8) if want_u8 else None
        _check(lib().gl_full_result(self.h, L.h, z.ctypes.data, z8.ctypes.data if want_u8 else None))
        z = z[:, :, 0] if ch == 1 else z
        return (z, z8[:, :, 0] if ch == 1 else z8) if want_u8 else z

    def upload(self, kind, a) -> Mat:
        a = np.ascontiguousarray(a, dtype=np.float64)
        rows, cols = (a.shape[0], 1) if a.ndim == 1 else a.shape
        o = C.c_void_p()
        _check(lib().gl_mat_upload(self.h, kind, a.ctypes.data, rows, cols, C.byref(o)))
        return Mat(self, o)

    def filter(self, phi: Mat, f_eigvals: Mat, gain=3.0, clip_low=False, want_u8=False):
        H, W, ch = self.shape
        z = np.zeros((H, W, ch), dtype=np.float32)
        z8 = np.zeros((H, W, ch), dtype=np.uint8) if want_u8 else None
        _check(lib().gl_filter(self.h, phi.h, f_eigvals.h, gain, int(clip_low), z.ctypes.data,
                               z8.ctypes.data if want_u8 else None))
        z = z[:, :, 0] if ch == 1 else z
        if want_u8:
            return z, (z8[:, :, 0] if ch == 1 else z8)
        return z

    def filter_resident(self, phi: Mat, f_eigvals: Mat, gain=3.0, clip_low=False):
        """gl_filter with the result left on the device (no host copy): for timing the stage calls."""
        _check(lib().gl_filter(self.h, phi.h, f_eigvals.h, gain, int(clip_low), None, None))

    # ---- whole path -------------------------------------------------------------------------
    def run(self, img, params: Params, z_out=None, z8_out=None, want_eigvals=True):
        """gl_run: H2D of `img`, all stages, D2H of z into z_out (float32 [H,W(,C)])."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        H, W = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        self.shape = (H, W, ch)
        if z_out is None:
            z_out = np.zeros(img.shape, dtype=np.float32)
        elif z_out is False:        # no fp32 result wanted (u8 only)
            z_out = None
        p, m = C.c_uint(), C.c_int()
        # the solver serves at most 8192 samples (uniform sampling may return up to ~4x the requested count); the library checks
        # the capacity it is given
        mu = np.zeros(min(H * W, 8192), dtype=np.float64) if want_eigvals else None
        _check(lib().gl_run(self.h, img.ctypes.data, W, H, ch, C.byref(params), z_out.ctypes.data if z_out is not None else None,
                            z8_out.ctypes.data if z8_out is not None else None, C.byref(p), C.byref(m),
                            mu.ctypes.data if want_eigvals else None, mu.size if want_eigvals else 0))
        return dict(z=z_out, p=p.value, m=m.value, mu=mu[:m.value] if want_eigvals else None)

    def run_resident(self, params: Params, z_out=None, want_eigvals=False):
        p, m = C.c_uint(), C.c_int()
        mu = np.zeros(8192, dtype=np.float64) if want_eigvals else None
        _check(lib().gl_run_resident(self.h, C.byref(params), z_out.ctypes.data if z_out is not None else None, None,
                                     C.byref(p), C.byref(m), mu.ctypes.data if want_eigvals else None, mu.size if want_eigvals else 0))
        return dict(p=p.value, m=m.value, mu=mu[:m.value] if want_eigvals else None)


# ---------------------------------------------------------------------------------------------
# Reference-shaped entry point: python/image_processing.py:244-357
# ---------------------------------------------------------------------------------------------
def image_processing(y, cr=None, cb=None, ctx: Context | None = None, **kwargs):
    """Filter the luma plane `y` (u8 [H,W]) like the reference's image_processing(y, cr, cb, **kwargs):
    kwargs['sampling'] in {'spatially_uniform','random'}, kwargs['affinity'] in {'bilateral','photometric',
    'spatial','NLM'}; 1 % of the pixels are sampled unless kwargs['sample_size'] is given.  Cr/Cb pass through
    untouched (python/image_processing.py:411-424).  Returns (z, cr, cb) with z float32 [H,W].

    The filter is the C code's z = y + 3 Phi Lambda Phi^T y (hpc/display.c:64-73); pass gain=-1, power=...
    for other choices."""
    own = ctx is None
    if own:
        ctx = Context()
    try:
        prm = default_params(sampling=kwargs.get("sampling", SPATIALLY_UNIFORM), affinity=kwargs.get("affinity", BILATERAL))
        for k in ("sample_size", "seed", "num_eigvals", "gain", "power", "gram_schmidt", "h_loc", "h_val", "clip_low"):
            if k in kwargs:
                setattr(prm, k, kwargs[k])
        r = ctx.run(y, prm, want_eigvals=False)
        return r["z"], cr, cb
    finally:
        if own:
            ctx.close()


# python/utils.py:33-44 (rgb2ycc / ycc2rgb: the YUV matrices of scikit-image)
YUV_FROM_RGB = np.array([[0.299, 0.587, 0.114],
                         [-0.14714119, -0.28886916, 0.43601035],
                         [0.61497538, -0.51496512, -0.10001026]])
RGB_FROM_YUV = np.linalg.inv(YUV_FROM_RGB)


def rgb2ycc(im):
    return np.dot(im, YUV_FROM_RGB.T)


def ycc2rgb(im):
    return np.dot(im, RGB_FROM_YUV.T)


def image_processing_rgb(rgb, ctx: Context | None = None, **kwargs):
    """The colour branch of the prototype's main (python/image_processing.py:411-424): RGB -> YCC, filter the luma plane
    only, pass Cr/Cb through, back to RGB.  The device path works on 8-bit samples, so the luma is rounded to u8 before
    filtering (the prototype filters the float luma); boundary plumbing on the host, the filter itself on the GPU.
    Returns float64 [H, W, 3] (not clipped, like the prototype before its uint8 cast)."""
    rgb = np.asarray(rgb)
    ycc = rgb2ycc(rgb.astype(np.float64))
    y8 = np.clip(np.rint(ycc[:, :, 0]), 0, 255).astype(np.uint8)
    z_y, _, _ = image_processing(y8, ycc[:, :, 1], ycc[:, :, 2], ctx=ctx, **kwargs)
    out = ycc.copy()
    out[:, :, 0] = ycc[:, :, 0] + (z_y.astype(np.float64) - y8)     # the filter's change, applied to the unrounded luma
    return ycc2rgb(out)


sampling_methods = {RANDOM: RANDOM, SPATIALLY_UNIFORM: SPATIALLY_UNIFORM}
affinity_methods = {SPATIAL: SPATIAL, NLM: NLM, BILATERAL: BILATERAL, PHOTOMETRIC: PHOTOMETRIC}
