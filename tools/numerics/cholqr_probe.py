"""CPU experiment (numpy/torch): is a column-scaled CholeskyQR2 a safe replacement for Householder QR inside the
randomised range finder of libctmb?  Emulates rsvd_batch (move.cu) on matrices M = R^T Rt taken from oracle CTM runs."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'oracle'))
import numpy as np, torch
import ctm_oracle as orc

def householder(Y):
    Q, R = np.linalg.qr(Y)
    return Q, R

def chol_reg(Gs, thr, stats):
    """right-looking Cholesky with per-pivot regularisation: a pivot below thr (scaled units) is raised to thr"""
    A = Gs.copy(); n = len(A); L = np.zeros_like(A); nreg = 0; minp = 1.0
    for j in range(n):
        p = A[j, j]
        minp = min(minp, p)
        if p < thr: p = thr; nreg += 1
        L[j, j] = np.sqrt(p)
        L[j+1:, j] = A[j+1:, j] / L[j, j]
        A[j+1:, j+1:] -= np.outer(L[j+1:, j], L[j+1:, j])
    stats.append(('minpivot', minp, nreg))
    return L

def scholqr2(Y, stats):
    k = Y.shape[1]
    R_tot = np.eye(k)
    Q = Y.copy()
    for p in range(2):
        G = Q.T @ Q
        d = np.sqrt(np.diag(G))
        junk = d <= JUNK * d.max()
        if not np.isfinite(d).all(): raise RuntimeError('nan')
        d[junk] = 1.0
        Gs = G / np.outer(d, d)
        Gs[junk, :] = 0; Gs[:, junk] = 0; Gs[junk, junk] = 1.0
        Q[:, junk] = 0
        if p == 1: stats.append(np.abs(Gs - np.eye(len(d))).max())
        L = chol_reg(Gs, THR, stats)
        R = L.T * d[None, :]
        Q = np.linalg.solve(R.T, Q.T).T
        R_tot = R @ R_tot
        stats.append(('junk', int(junk.sum())))
    return Q, R_tot

JUNK = 0.0
THR = 1e-10
def rsvd(M, chi, k, q, plan, seed=0):
    """plan: string of 'H'/'C' per QR (q+2 QRs: Y0..Yq, Z)"""
    rng = np.random.default_rng(seed)
    n = M.shape[0]
    Om = rng.standard_normal((n, k))
    stats = []
    def qr(Y, i):
        return householder(Y) if plan[i] == 'H' else scholqr2(Y, stats)
    Q, _ = qr(M @ Om, 0)
    for it in range(q):
        Z = M.T @ Q
        Q, _ = qr(M @ Z, it + 1)
    Z = M.T @ Q
    Q2, R2 = qr(Z, q + 1)
    # Jacobi stand-in: LAPACK on the small factor (the question here is the QR, not the small SVD)
    Uh, s, Wh = np.linalg.svd(R2.T)
    U = Q @ Uh; V = Q2 @ Wh.T
    res = np.abs(M @ V[:, :chi] - U[:, :chi] * s[:chi]).max() / s[0]
    orthU = np.abs(U[:, :chi].T @ U[:, :chi] - np.eye(chi)).max()
    return U[:, :chi], s[:chi + 1], V[:, :chi], res, orthU, stats

def matrices(D, chi, family, iters):
    sites = orc.random_state_4site(D, family=family)
    C, T = orc.init_env(sites, orc.v2s_4site, chi)
    out = []
    for it in range(iters):
        for d in orc.DIRECTIONS:
            for rep in range(2):
                R, Rt = orc.halves(d, (0, 0), sites, orc.v2s_4site, C, T)
                out.append((R.numpy(), Rt.numpy()))
                orc.ctm_move(d, sites, orc.v2s_4site, C, T, chi)
    return out

if __name__ == '__main__':
    D, chi = int(sys.argv[1]), int(sys.argv[2])
    k = int(np.ceil(1.75 * chi)); q = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    for fam in 'AB':
        ms = matrices(D, chi, fam, 2)
        for plan in ('H' * (q + 2), 'HH' + 'C' * q, 'HH' + 'C' * (q - 1) + 'H', 'HH' + 'C' * (q - 2) + 'HH'):
            worst = dict(res=0, orth=0, ds=0, dP=0, e=0)
            flags = 0
            for (R, Rt) in ms[::3]:
                M = R.T @ Rt
                U, s, V, res, orthU, stats = rsvd(M, chi, k, q, plan)
                sr = np.linalg.svd(M, compute_uv=False)
                keep = sr[:chi] / sr[0] > 1e-8
                ds = np.abs(s[:chi][keep] - sr[:chi][keep]).max() / sr[0]
                dsr = (np.abs(s[:chi][keep] - sr[:chi][keep]) / sr[:chi][keep]).max()
                # gauge-invariant projector product
                sq = np.where(keep, 1 / np.sqrt(np.where(keep, s[:chi], 1)), 0)
                P = (R @ U) * sq; Pt = (Rt @ V) * sq
                Ur, srr, Vhr = np.linalg.svd(M)
                sqr = np.where(keep, 1 / np.sqrt(np.where(keep, srr[:chi], 1)), 0)
                Pr = (R @ Ur[:, :chi]) * sqr; Ptr = (Rt @ Vhr[:chi].T) * sqr
                dP = np.abs(P @ Pt.T - Pr @ Ptr.T).max() / np.abs(Pr @ Ptr.T).max()
                flags += sum(x[2] for x in stats if isinstance(x, tuple) and x[0] == 'minpivot')
                worst['minp'] = min(worst.get('minp', 1), min([x[1] for x in stats if isinstance(x, tuple) and x[0] == 'minpivot'], default=1))
                worst['junk'] = max(worst.get('junk', 0), max([x[1] for x in stats if isinstance(x, tuple) and x[0] == 'junk'], default=0))
                e = max([x for x in stats if not isinstance(x, tuple)], default=0)
                worst.update(res=max(worst['res'], res), orth=max(worst['orth'], orthU), ds=max(worst['ds'], dsr), dP=max(worst['dP'], dP), e=max(worst['e'], e))
            print(f'fam {fam} plan {plan}: resid {worst["res"]:.2e} orthU {worst["orth"]:.1e} rel dS {worst["ds"]:.1e} d(P Pt^T) {worst["dP"]:.1e} '
                  f'pass-2 |G-I| {worst["e"]:.1e} regularised pivots {flags} min pivot {worst.get("minp",1):.1e} junk cols {worst.get("junk",0)}', flush=True)
