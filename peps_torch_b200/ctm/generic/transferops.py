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


def _ring_mv(Ts, V):
    """The ring MPO of the tensors ``Ts[i][chi_i, chi_{i+1}, o_i, v_i]`` (chi_L = chi_0: periodic) applied to ``V[v_0 .. v_{L-1}]``
    -- the mat-vec of transferops.py:304-318 as L pairwise contractions through libctmb."""
    from ... import ad
    eng = corrf._engine()
    L = len(Ts)
    v = [chr(ord('a') + i) for i in range(L)]
    o = [chr(ord('A') + i) for i in range(L)]
    run = 'y'
    idx = ''.join(v[1:]) + 'x' + run + o[0]
    cur = ad.contract(eng, f'x{run}{o[0]}{v[0]},{"".join(v)}->{idx}', Ts[0], V)
    for i in range(1, L - 1):
        nxt = 'z' if run == 'y' else 'y'
        new = ''.join(ch for ch in idx if ch not in (run, v[i])) + nxt + o[i]
        cur = ad.contract(eng, f'{run}{nxt}{o[i]}{v[i]},{idx}->{new}', Ts[i], cur)
        idx, run = new, nxt
    return ad.contract(eng, f'{run}x{o[L - 1]}{v[L - 1]},{idx}->{"".join(o)}', Ts[L - 1], cur)


def _ttensor(state, env, c, d):
    """T of site ``c`` on side ``d`` as [chi, chi, D, D] (ket, bra), the two chi legs in the order transferops.py:285-298 fixes."""
    leg = {(0, -1): 1, (-1, 0): 2, (0, 1): 3, (1, 0): 4}[d]
    perm = {(0, -1): (0, 2, 1), (-1, 0): (0, 1, 2), (0, 1): (1, 2, 0), (1, 0): (0, 2, 1)}[d]
    D = state.site(c).size(leg)
    return env.T[(c, d)].permute(*perm).contiguous().view(env.chi, env.chi, D, D)


def get_EH_spec_Ttensor(n, L, coord, direction, state, env, verbosity=0):
    r"""Leading ``n`` eigenvalues of the approximate :math:`\exp(-H_{ent})` of an L-leg cylinder: the product of the two ring MPOs
    built from the T tensors on side ``direction`` and on the opposite side (transferops.py:207-370; one-site unit cells)."""
    import warnings
    assert L > 1, "L must be larger than 1"
    assert state.lX == state.lY == 1, "only single-site unit cell is supported"
    dir_to_ind = {(0, -1): 1, (-1, 0): 2, (0, 1): 3, (1, 0): 4}
    d_grow = {(0, -1): (1, 0), (-1, 0): (0, -1), (0, 1): (-1, 0), (1, 0): (0, 1)}[direction]
    d_opp = (-direction[0], -direction[1])
    cs = [state.vertexToSite((coord[0] + i * d_grow[0], coord[1] + i * d_grow[1])) for i in range(L)]
    ads = [state.site(c).size(dir_to_ind[direction]) for c in cs]
    dim = int(np.prod(ads))
    if dim <= n:
        warnings.warn("Total dimension of H_ent operator is <= n.", RuntimeWarning)
        return None
    device, dtype = _dev_dtype(state, env)
    rings = [[_ttensor(state, env, c, d) for c in cs] for d in (direction, d_opp)]

    def _mv(v0):
        V = torch.as_tensor(v0).to(dtype=dtype, device=device).view(ads)
        for Ts in rings:
            V = _ring_mv(Ts, V.contiguous())
        return V.reshape(dim).cpu().numpy()
    with torch.no_grad():
        op = LinearOperator((dim, dim), matvec=_mv, dtype="complex128" if dtype.is_complex else "float64")
        vals = np.copy(eigs(op, k=n, v0=None, return_eigenvectors=False)[::-1])
    vals = (1.0 / np.abs(vals[0])) * vals
    S = torch.zeros((n, 2), dtype=torch.float64, device=next(iter(state.sites.values())).device)
    S[:, 0] = torch.as_tensor(np.real(vals))
    S[:, 1] = torch.as_tensor(np.imag(vals))
    return S
