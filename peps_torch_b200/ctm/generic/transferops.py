"""Leading spectrum of the transfer operators (ctm/generic/transferops.py:14-205 of peps-torch: get_Top_w0_spec, get_Top_spec),
the tail of every ctmrg_*.py script.  Same structure as the reference -- scipy's ARPACK drives a LinearOperator -- with the
mat-vec running on the GPU through libctmb's contraction chains (ctm/generic/corrf.py of this package); per mat-vec one vector
travels host -> device and back, which is what ARPACK's reverse-communication interface asks for."""
import numpy as np
import torch
from scipy.sparse.linalg import LinearOperator, eigs
from . import corrf


def _get_chis(state, env, coord, direction, width):
    assert direction in [(0, -1), (0, 1), (-1, 0), (1, 0)], "Invalid direction: " + str(direction)
    coord = state.vertexToSite(coord)
    if direction in [(0, -1), (0, 1)]:
        cs = state.vertexToSite((coord[0] + width, coord[1]))
        if direction == (0, -1):
            return env.T[(coord, (-1, 0))].size(1), env.T[(cs, (1, 0))].size(2)
        return env.T[(coord, (-1, 0))].size(0), env.T[(cs, (1, 0))].size(0)
    cs = state.vertexToSite((coord[0], coord[1] + width))
    if direction == (-1, 0):
        return env.T[(coord, (0, -1))].size(2), env.T[(cs, (0, 1))].size(2)
    return env.T[(coord, (0, -1))].size(0), env.T[(cs, (0, 1))].size(1)


def _leading(n, dim, mv, cplx, device, eigenvectors=False, normalize=True):
    T = LinearOperator((dim, dim), matvec=mv, dtype="complex128" if cplx else "float64")
    if eigenvectors:
        vals, vecs = eigs(T, k=n, v0=None, return_eigenvectors=True)
    else:
        vals = eigs(T, k=n, v0=None, return_eigenvectors=False)
    ind = np.argsort(np.abs(vals))[::-1]
    vals = vals[ind]
    if normalize:
        vals = (1.0 / np.abs(vals[0])) * vals
    L = torch.zeros((n, 2), dtype=torch.float64, device=device)
    L[:, 0] = torch.as_tensor(np.real(vals))
    L[:, 1] = torch.as_tensor(np.imag(vals))
    if eigenvectors:
        return L, torch.as_tensor(vecs[:, ind], device=device)
    return L


def _steps(state, direction):
    if direction in [(1, 0), (-1, 0)]:
        return state.lX
    if direction in [(0, 1), (0, -1)]:
        return state.lY
    raise ValueError("Invalid direction: " + str(direction))


def _dev_dtype(state, env):
    t = next(iter(env.T.values()))
    return t.device, t.dtype


def get_Top_w0_spec(n, coord, direction, state, env, verbosity=0):
    r"""Leading ``n`` eigenvalues of the width-0 transfer operator (transferops.py:38-108), normalised by the largest
    magnitude, as an ``n x 2`` tensor (real, imaginary part)."""
    chi1, chi2 = _get_chis(state, env, coord, direction, 0)
    N = _steps(state, direction)
    device, dtype = _dev_dtype(state, env)

    def _mv(v):
        c0 = coord
        V = torch.as_tensor(v).to(dtype=dtype, device=device).view(chi1, chi2)
        for _ in range(N):
            V = corrf.apply_TM_0sO(c0, direction, state, env, V, verbosity=verbosity)
            c0 = (c0[0] + direction[0], c0[1] + direction[1])
        return V.reshape(chi1 * chi2).cpu().numpy()
    with torch.no_grad():
        return _leading(n, chi1 * chi2, _mv, dtype.is_complex, device)


def get_Top_spec(n, coord, direction, state, env, eigenvectors=False, verbosity=0):
    r"""Leading ``n`` eigenvalues (optionally eigenvectors) of the width-1 transfer operator (transferops.py:110-205)."""
    dir_to_ind = {(0, -1): 1, (-1, 0): 2, (0, 1): 3, (1, 0): 4}
    chi1, chi2 = _get_chis(state, env, coord, direction, 0)
    a = state.site(coord)
    ad_ = a.size(dir_to_ind[(-direction[0], -direction[1])])
    d2 = ad_ * ad_ if a.dim() == 5 else a.size(dir_to_ind[(-direction[0], -direction[1])] - 1)
    N = _steps(state, direction)
    device, dtype = _dev_dtype(state, env)

    def _mv(v):
        c0 = coord
        V = torch.as_tensor(v).to(dtype=dtype, device=device).view(chi1, d2, chi2)
        for _ in range(N):
            V = corrf.apply_TM_1sO(c0, direction, state, env, V, verbosity=verbosity)
            c0 = (c0[0] + direction[0], c0[1] + direction[1])
        return V.reshape(chi1 * d2 * chi2).cpu().numpy()
    with torch.no_grad():
        return _leading(n, chi1 * d2 * chi2, _mv, dtype.is_complex, device, eigenvectors)
