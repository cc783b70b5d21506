"""First-light script for the GPU box: prints an error table stage by stage."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import ctm_oracle as orc
import helpers as H
from peps_torch_b200.engine import CtmEngine

dev = torch.device('cuda:0')
eng = CtmEngine()
torch.manual_seed(0)

def rel(a, b): return H.maxrel(a.cpu(), b.cpu())

print('== einsum2')
for dt in (torch.float64, torch.complex128):
    for spec, sa, sb in [('ab,buc->auc', (7, 5), (5, 3, 11)), ('auc,ael->ucel', (13, 4, 9), (13, 9, 4)),
                         ('ik,kj->ij', (130, 70), (70, 150)), ('ki,kj->ij', (200, 129), (200, 65)),
                         ('pqcert,sprfg->qcetsfg', (3, 3, 5, 5, 3, 3), (2, 3, 3, 3, 3))]:
        A = torch.randn(sa, dtype=dt, device=dev); B = torch.randn(sb, dtype=dt, device=dev)
        for ca in (False, True):
            out = eng.einsum2(spec, A, B, conjA=ca, conjB=False)
            ref = torch.einsum(spec, A.conj() if ca else A, B)
            print(f'  {str(dt):18s} {spec:28s} conjA={ca} err={rel(out, ref):.2e}')

for name in ['generic_4site_D2_chi8_B', 'generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A']:
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'mid_')
    coord = list(sites.keys())[-1]
    print('==', name)
    for kind in orc.CORNERS:
        kc, k1, k2, _ = orc.CORNERS[kind]
        out = eng.c2x2(kind, C[(coord, kc)].to(dev), T[(coord, k1)].to(dev), T[(coord, k2)].to(dev), sites[coord].to(dev), chi)
        print(f'  c2x2_{kind} err={rel(out, torch.from_numpy(z["c2x2_" + kind])):.2e}')
    for d in orc.DIRECTIONS:
        tg = f'{d[0]}_{d[1]}'
        R, Rt = torch.from_numpy(z[f'halves_{tg}_R']), torch.from_numpy(z[f'halves_{tg}_Rt'])
        Pr, Ptr = torch.from_numpy(z[f'proj_{tg}_P']), torch.from_numpy(z[f'proj_{tg}_Pt'])
        M = (R.t() @ Rt)
        U, S, V = eng.truncated_svd(M.to(dev), chi)
        Sr = torch.linalg.svdvals(M)[:chi]
        rec = (U * S.to(U.dtype)) @ V.conj().t()
        Ur, Srf, Vhr = torch.linalg.svd(M)
        best = (Ur[:, :chi] * Srf[:chi].to(Ur.dtype)) @ Vhr[:chi]
        print(f'  svd{d}: S err={float((S.cpu() - Sr).abs().max() / Sr[0]):.2e} rank-chi approx err={rel(rec, best):.2e} '
              f'orthU={float((U.conj().t() @ U - torch.eye(chi, device=dev)).abs().max()):.2e}')
        P, Pt, S2 = eng.projectors(R.to(dev), Rt.to(dev), chi)
        print(f'  proj{d}: P.Pt^T err={rel(P @ Pt.t(), Pr @ Ptr.t()):.2e}  |P| err={rel(P.abs(), Pr.abs()):.2e}')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    for d in orc.DIRECTIONS:
        env = H.Env(chi, H.to_dev(C, dev), H.to_dev(T, dev))
        eng.move_generic(d, st, env)
        Cg, Tg = H.golden_env(z, f'move_{d[0]}_{d[1]}_')
        print(f'  move{d}: |C|,|T| err={H.env_abs_diff(env.C, env.T, Cg, Tg):.2e}')
    # full run from the initial env
    C0, T0 = H.golden_env(z, 'init_')
    env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
    n_iter = meta['n_iter'] + 1
    for it in range(n_iter):
        for d in orc.DIRECTIONS:
            for _ in range(lX if d in (orc.LEFT, orc.RIGHT) else lY):
                eng.move_generic(d, st, env)
    Cf, Tf = H.golden_env(z, 'final_')
    print(f'  run {n_iter} iters: |C|,|T| err={H.env_abs_diff(env.C, env.T, Cf, Tf):.2e} spectra err={H.spectra_diff(env.C, Cf):.2e}')
    if 'energy' in z.files:
        e = orc.energy_j1j2(sites, v2s, {k: v.cpu() for k, v in env.C.items()}, {k: v.cpu() for k, v in env.T.items()}, 1.0, meta['j2'])
        print(f'  energy {e:.14f} ref {float(z["energy"][0]):.14f} diff {abs(e - float(z["energy"][0])):.2e}')

for name in ['c4v_D2_chi8_A', 'c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128']:
    z, meta = H.load_golden(name)
    chi = meta['chi']
    a = torch.from_numpy(z['site']).to(dev)
    print('==', name)
    Cm, Tm = torch.from_numpy(z['mid_C']).to(dev), torch.from_numpy(z['mid_T']).to(dev)
    M = torch.from_numpy(z['mid_c2x2'])
    D, U = eng.truncated_eig_sym(M.to(dev), chi)
    Dr, Ur = orc.truncated_eig_sym(M, chi)
    print(f'  eig: D err={float((D.cpu() - Dr).abs().max()):.2e} resid={float((M.to(dev) @ U - U * D.to(U.dtype)).abs().max()):.2e}')
    nC, nT, Dv = eng.move_c4v(a, Cm, Tm, chi)
    print(f'  move: C err={rel(nC, torch.from_numpy(z["mid_nC"])):.2e} |T| err={rel(nT.abs(), torch.from_numpy(z["mid_nT"]).abs()):.2e}')
    Cc, Tc = torch.from_numpy(z['init_C']).to(dev), torch.from_numpy(z['init_T']).to(dev)
    for i in range(meta['n_iter']):
        Cc, Tc, _ = eng.move_c4v(a, Cc, Tc, chi)
    print(f'  {meta["n_iter"]} moves: C err={rel(Cc, torch.from_numpy(z["final_C"])):.2e} |T| err={rel(Tc.abs(), torch.from_numpy(z["final_T"]).abs()):.2e}')
    e = orc.energy_j1j2_c4v(a.cpu(), Cc.cpu(), Tc.cpu(), 1.0, meta['j2'])
    print(f'  energy {e:.14f} ref {float(z["energy"][0]):.14f}')
torch.cuda.synchronize()
print('launches, flops', eng.counters())
