"""Leading spectrum of the width-1 and width-2 transfer operators of the C4v ansatz (ctm/one_site_c4v/transferops_c4v.py of
peps-torch: get_Top_spec_c4v :10-68, get_Top2_spec_c4v :70-117): scipy's ARPACK over a LinearOperator whose mat-vec is one
libctmb chain (corrf_c4v.apply_TM_1sO / apply_TM_1sO_2)."""
import numpy as np
import torch
from scipy.sparse.linalg import LinearOperator, eigs
from ..generic.transferops import _leading
from . import corrf_c4v


def get_Top_spec_c4v(n, state, env_c4v, normalize=True, eigenvectors=False, verbosity=0):
    r"""Leading ``n`` eigenvalues (``n x 2``: real, imaginary part; divided by the largest magnitude if ``normalize``), and the
    eigenvectors if asked for."""
    a = next(iter(state.sites.values()))
    T = env_c4v.T[env_c4v.keyT]
    chi, d2 = T.size(0), T.size(2)
    device, dtype = T.device, T.dtype

    def _mv(v):
        V = torch.as_tensor(v).to(dtype=dtype, device=device).view(chi, d2, chi)
        return corrf_c4v.apply_TM_1sO(state, env_c4v, V, verbosity=verbosity).reshape(chi * d2 * chi).cpu().numpy()
    with torch.no_grad():
        return _leading(n, chi * d2 * chi, _mv, dtype.is_complex, a.device, eigenvectors, normalize)


def get_Top2_spec_c4v(n, state, env_c4v, verbosity=0):
    r"""Leading ``n`` eigenvalues of the width-2 transfer operator (``n x 2``: real, imaginary part), in ARPACK's order reversed
    and divided by the magnitude of the first, as transferops_c4v.py:105-116 returns them."""
    a = next(iter(state.sites.values()))
    T = env_c4v.T[env_c4v.keyT]
    chi, d2 = T.size(0), T.size(2)
    device, dtype = T.device, T.dtype
    dim = chi * d2 * d2 * chi

    def _mv(v):
        V = torch.as_tensor(v).to(dtype=dtype, device=device).view(chi, d2, d2, chi)
        return corrf_c4v.apply_TM_1sO_2(state, env_c4v, V, verbosity=verbosity).reshape(dim).cpu().numpy()
    with torch.no_grad():
        op = LinearOperator((dim, dim), matvec=_mv, dtype="complex128" if dtype.is_complex else "float64")
        vals = np.copy(eigs(op, k=n, v0=None, return_eigenvectors=False)[::-1])
    vals = (1.0 / np.abs(vals[0])) * vals
    L = torch.zeros((n, 2), dtype=torch.float64, device=a.device)
    L[:, 0] = torch.as_tensor(np.real(vals))
    L[:, 1] = torch.as_tensor(np.imag(vals))
    return L


def get_EH_spec_Ttensor(n, L, state, env_c4v, verbosity=0):
    r"""Leading ``n`` eigenvalues of the approximate :math:`\exp(-H_{ent})` of an L-leg cylinder, the ring MPO of L copies of
    the C4v T tensor (transferops_c4v.py:119-181)."""
    from ..generic.transferops import _ring_mv
    assert L > 1, "L must be larger than 1"
    a = next(iter(state.sites.values()))
    T0 = env_c4v.T[env_c4v.keyT]
    chi, D = T0.size(0), a.size(4)
    T = T0.contiguous().view(chi, chi, D, D)
    dim = D ** L

    def _mv(v0):
        V = torch.as_tensor(v0).to(dtype=T.dtype, device=T.device).view([D] * L)
        return _ring_mv([T] * L, V.contiguous()).reshape(dim).cpu().numpy()
    with torch.no_grad():
        op = LinearOperator((dim, dim), matvec=_mv, dtype="complex128" if T.dtype.is_complex else "float64")
        vals = np.copy(eigs(op, k=n, v0=None, return_eigenvectors=False)[::-1])
    vals = (1.0 / np.abs(vals[0])) * vals
    S = torch.zeros((n, 2), dtype=torch.float64, device=a.device)
    S[:, 0] = torch.as_tensor(np.real(vals))
    S[:, 1] = torch.as_tensor(np.imag(vals))
    return S
