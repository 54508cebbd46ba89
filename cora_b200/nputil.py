"""numpy-facing helpers of the path (mirrors ``cora/util/nputil.py:51-125``), computed on the GPU."""

import ctypes

import numpy as np

from . import _dev, _lib


def root_batched_device(cl_dev, jitter_rel=0.0, clip_rel=1e-16, stream=None, out=None, ws=None):
    """Batched root on device: ``cl_dev`` CUDA float64 [nl, nz, nz] -> (root, used_eigh, num_pos).

    Semantics of ``skysim.py:116-119`` + ``nputil.py:51-101`` (``truncate=False``): jitter on the
    diagonal, Cholesky, eigen-decomposition with clipping where Cholesky meets a non-positive pivot.
    """
    t = _dev.torch()
    nl, nz = int(cl_dev.shape[0]), int(cl_dev.shape[1])
    if out is None:
        out = (_dev.empty((nl, nz, nz), t.float64), _dev.empty((nl,), t.int32), _dev.empty((nl,), t.int32))
    root, used, npos = out
    if ws is None:
        ws = root_workspace(nl, nz)
    nbytes = ws.numel()
    _lib.call("cora_b200_root_batched", _lib.ptr(cl_dev), nl, nz, float(jitter_rel), float(clip_rel), _lib.ptr(root),
              _lib.ptr(used), _lib.ptr(npos), _lib.ptr(ws), int(nbytes), _lib.stream_ptr(stream))
    return root, used, npos


def root_multi_workspace(nblocks, nl, nz, max_eigh=None):
    """Workspace for ``root_batched_multi_device`` (room for ``max_eigh`` l's on the eigen branch at once)."""
    lib = _lib.load()
    full = lib.cora_b200_root_multi_workspace_bytes(nblocks, nl, nz)
    per_l = 16 * nz * nz * nblocks
    one = full - per_l * (nl - 1)
    want = full if max_eigh is None else one + per_l * (max(1, min(nl, max_eigh)) - 1)
    return _dev.workspace(min(want, max(one, _dev.free_bytes() - (2 << 30))))


def root_batched_multi_device(cl_blocks, jitter_rel=1e-14, clip_rel=1e-16, stream=None, out=None, ws=None, zero_scale=None):
    """Root of ``blockdiag(cl_blocks[0][l], cl_blocks[1][l], ..., 0)`` kept as separate blocks, with the
    reference's whole-matrix semantics (``cora/scripts/makesky.py:368-382`` + ``skysim.py:116-119`` +
    ``nputil.py:81-96``): one jitter, one Cholesky-or-eigh decision and one eigenvalue clip threshold per l
    (``cora_b200_root_batched_multi``).  ``cl_blocks``: list of CUDA float64 ``[nl, nz, nz]``.
    Returns ``(roots list, used int32[nblocks, nl], num_pos int32[nblocks, nl])``; ``zero_scale`` (CUDA
    float64[nl], optional) receives the root scale of an implicit all-zero block (Stokes V)."""
    t = _dev.torch()
    nb = len(cl_blocks)
    nl, nz = int(cl_blocks[0].shape[0]), int(cl_blocks[0].shape[1])
    if out is None:
        out = ([_dev.empty((nl, nz, nz), t.float64) for _ in range(nb)], _dev.empty((nb, nl), t.int32),
               _dev.empty((nb, nl), t.int32))
    roots, used, npos = out
    if ws is None:
        ws = root_multi_workspace(nb, nl, nz)
    cls = (ctypes.c_void_p * nb)(*[c.data_ptr() for c in cl_blocks])
    rts = (ctypes.c_void_p * nb)(*[r.data_ptr() for r in roots])
    _lib.call("cora_b200_root_batched_multi", cls, nb, nl, nz, float(jitter_rel), float(clip_rel), rts, _lib.ptr(used),
              _lib.ptr(npos), _lib.ptr(zero_scale), _lib.ptr(ws), int(ws.numel()), _lib.stream_ptr(stream))
    return roots, used, npos


def eigh_batched_device(a_dev):
    """Batched ``scipy.linalg.eigh`` on device: CUDA float64 ``[nl, nz, nz]`` -> ``(evals[nl, nz]``
    ascending, ``evecs[nl, nz, nz]`` with eigenvectors in columns``)`` (``cora_b200_eigh_batched``)."""
    t = _dev.torch()
    lib = _lib.load()
    nl, nz = int(a_dev.shape[0]), int(a_dev.shape[1])
    evecs = _dev.empty((nl, nz, nz), t.float64)
    evals = _dev.empty((nl, nz), t.float64)
    full = lib.cora_b200_eigh_workspace_bytes(nl, nz)
    one = lib.cora_b200_eigh_workspace_bytes(1, nz) + 32 * nl * (nz + 4)
    ws = _dev.workspace(min(full, max(one, _dev.free_bytes() - (2 << 30))))
    _lib.call("cora_b200_eigh_batched", _lib.ptr(a_dev), nl, nz, _lib.ptr(evecs), _lib.ptr(evals), _lib.ptr(ws), int(ws.numel()),
              _lib.stream_ptr())
    return evals, evecs


def root_workspace(nl, nz, max_eigh=None):
    """Workspace for ``root_batched_device``: room for ``max_eigh`` simultaneous eigen fallbacks
    (default: all nl matrices, bounded by free memory)."""
    lib = _lib.load()
    full = lib.cora_b200_root_workspace_bytes(nl, nz)
    one = full - 16 * nz * nz * (nl - 1)
    want = full if max_eigh is None else one + 16 * nz * nz * (max(1, min(nl, max_eigh)) - 1)
    return _dev.workspace(min(want, max(one, _dev.free_bytes() - (2 << 30))))


def matrix_root_manynull(mat, threshold=1e-16, truncate=True):
    """Square root a matrix: Cholesky, else eigen-decomposition with small/negative eigenvalues
    set to zero (``nputil.py:51-101``).  Returns ``root`` or ``(root, num_pos)`` if ``truncate``.

    The eigen branch returns columns in ascending-eigenvalue order; with ``truncate`` only the
    last ``num_pos`` columns are kept and -- reference quirk preserved -- the array then has
    shape ``(1, N, num_pos)`` (``nputil.py:92-96``).

    The eigen branch has the reference's semantics at every size: eigenvalues below ``threshold`` times the
    largest are dropped, the columns are ``evecs * sqrt(evals)`` in ascending order.  Matrices larger than
    128 x 128 get there through a low-rank factorisation when they are numerically rank-deficient positive
    semi-definite (``csrc/root.cu``: pivoted Cholesky + one-sided Jacobi on the factor: the same eigenpairs, to
    high relative accuracy), otherwise through the full Jacobi sweep.  Eigenvalues within a few ulp of
    ``threshold * max`` are round-off in any implementation (LAPACK's included): ``num_pos`` can differ from
    scipy's by the number of eigenvalues sitting in that band.
    """
    t = _dev.torch()
    mat = np.asarray(mat, dtype=np.float64)
    cl = _dev.to_device(mat[np.newaxis], t.float64)
    root, used, npos = root_batched_device(cl, 0.0, threshold)
    r = root[0].cpu().numpy()
    eigh = bool(used[0].item())
    num_pos = int(npos[0].item()) if eigh else mat.shape[0]
    if not truncate:
        return r
    if eigh:
        r = r[:, -num_pos:][np.newaxis] if num_pos else r[np.newaxis]
    return r, num_pos


def complex_std_normal(shape, rng=None, seed=None):
    """Complex standard normal variates ``(N(0,1) + i N(0,1)) / sqrt(2)`` (``nputil.py:104-125``).

    With ``rng`` (a numpy Generator) the caller's stream is consumed exactly as the reference
    does -- real block, then imaginary block -- which is the identical-draw parity path.  With
    ``rng=None`` the variates come from the device Philox4x32-10 generator (``seed`` or a seed
    taken from numpy's legacy global state, which is what the reference would have consumed).
    """
    if rng is not None:
        return (rng.standard_normal(shape) + 1.0j * rng.standard_normal(shape)) / 2**0.5
    t = _dev.torch()
    shape = tuple(np.atleast_1d(shape).astype(int))
    n = int(np.prod(shape))
    if seed is None:
        seed = int(np.random.randint(0, 2**31 - 1))
    # one "l" = n - 1 with a 1 x 1 identity root, written as a single slab: out[m] = g[m], m = 0 .. n-1
    lmax = n - 1
    out = _dev.empty((n,), t.complex128)
    root = _dev.to_device(np.ones((1, 1, 1)), t.float64)
    llist = np.array([lmax], dtype=np.int32)
    row0 = np.zeros(1, dtype=np.int64)
    nu_base = _dev.zeros((1,), t.int64)
    nu_width = _dev.to_device(np.ones(1, dtype=np.int32), t.int32)
    nbytes = _lib.load().cora_b200_draw_apply_workspace_bytes(1, lmax, 1)
    ws = _dev.workspace(nbytes)
    _lib.call("cora_b200_draw_apply_slabs", _lib.ptr(root), _lib.ptr(llist), None, 1, 1, lmax, ctypes.c_ulonglong(seed), None, 0,
              _lib.ptr(row0), _lib.ptr(nu_base), _lib.ptr(nu_width), _lib.ptr(out), _lib.ptr(ws), int(nbytes), _lib.stream_ptr())
    return out.cpu().numpy().reshape(shape)
