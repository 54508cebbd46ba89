"""Oracle restatement of ``cora/util/nputil.py`` (test infrastructure, see package doc)."""

import numpy as np
import scipy.linalg as la


def matrix_root_manynull(mat, threshold=1e-16, truncate=True):
    """Cholesky-else-eigh-clip matrix root.

    Follows ``cora/util/nputil.py:51-101``: try ``scipy.linalg.cholesky(lower=True)``;
    on ``LinAlgError`` use ``eigh``, zero eigenvalues below ``max * threshold`` and return
    ``evecs * sqrt(evals)`` (columns in ascending-eigenvalue order; clipped columns are
    kept as zeros unless ``truncate``).
    """
    try:
        root = la.cholesky(mat, lower=True)
        num_pos = mat.shape[0]
    except la.LinAlgError:
        evals, evecs = la.eigh(mat)
        evals[evals < evals.max() * threshold] = 0.0
        num_pos = int(np.count_nonzero(evals))
        if truncate:
            # Reference quirk kept: ``evals[np.newaxis, -num_pos:]`` makes evals 2-D, so the
            # truncated eigh-branch root comes back with shape (1, N, num_pos) (nputil.py:92-96).
            evals = evals[np.newaxis, -num_pos:]
            evecs = evecs[:, -num_pos:]
        root = evecs * evals[np.newaxis, :] ** 0.5
    return (root, num_pos) if truncate else root


def complex_std_normal(shape, rng=None):
    """``(N(0,1) + i N(0,1)) / sqrt(2)``; real block drawn first, then the imaginary block.

    Follows ``cora/util/nputil.py:104-125`` (``rng=None`` uses numpy's legacy global state).
    """
    draw = np.random.standard_normal if rng is None else rng.standard_normal
    re = draw(shape)
    im = draw(shape)
    return (re + 1.0j * im) / 2**0.5
