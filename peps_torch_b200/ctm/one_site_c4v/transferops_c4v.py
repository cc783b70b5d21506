"""Leading spectrum of the width-1 transfer operator of the C4v ansatz (ctm/one_site_c4v/transferops_c4v.py:10-68 of
peps-torch, get_Top_spec_c4v): scipy's ARPACK over a LinearOperator whose mat-vec is one libctmb chain (corrf_c4v.apply_TM_1sO)."""
import torch
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
