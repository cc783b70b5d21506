"""Pin the oracle against the reference itself and write tests/golden/*.npz.

Run in the BUILD container only (needs /root/reference, which never travels):

    python oracle/gen_golden.py

For every function on the hot path the reference (imported unmodified from
/root/reference) and oracle/ctm_oracle.py are run on identical seeded inputs and
compared; the script aborts on the first mismatch.  The reference outputs are then
stored as small fixtures so that tests/ can re-check the oracle (CPU) and the CUDA
path (GPU box) without the reference being present.
"""
import os
import sys
import json
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
GOLD = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
os.chdir('/tmp')  # config.configure writes log files into cwd

import ctm_oracle as orc                                     # noqa: E402
import config as cfg                                          # noqa: E402
from ipeps.ipeps import IPEPS                                 # noqa: E402
from ipeps.ipeps_c4v import IPEPS_C4V                         # noqa: E402
from ctm.generic.env import ENV, init_env                     # noqa: E402
from ctm.generic import ctmrg, rdm                            # noqa: E402
from ctm.generic import ctm_components as comp                # noqa: E402
from ctm.generic.ctm_projectors import ctm_get_projectors_from_matrices  # noqa: E402
from ctm.one_site_c4v.env_c4v import ENV_C4V                  # noqa: E402
from ctm.one_site_c4v import env_c4v, ctmrg_c4v               # noqa: E402
from ctm.one_site_c4v.ctm_components_c4v import c2x2_sl       # noqa: E402
from linalg.custom_eig import truncated_eig_sym               # noqa: E402
from models import j1j2                                       # noqa: E402
from groups.pg import make_c4v_symm                           # noqa: E402

# the RDM dispatch patch of SURVEY 8c (opt_einsum is absent here)
rdm.rdm2x2 = lambda coord, state, env, **kw: rdm.rdm2x2_legacy(coord, state, env)


def set_dtype(dt):
    cfg.global_args.torch_dtype = dt
    cfg.global_args.dtype = 'complex128' if dt.is_complex else 'float64'
    cfg.global_args.device = 'cpu'


def maxrel(a, b):
    return float((a - b).abs().max() / max(b.abs().max(), 1e-300))


def check(name, a, b, tol):
    err = maxrel(a, b)
    print(f'  {name:34s} {err:.2e}')
    assert err <= tol, (name, err)


def ref_state_4site(sites):
    return IPEPS(sites, vertexToSite=orc.v2s_4site, lX=2, lY=2)


def dump_env(prefix, C, T, out):
    for (c, v), t in C.items():
        out[f'{prefix}C_{c[0]}{c[1]}_{v[0]}_{v[1]}'] = t.numpy()
    for (c, v), t in T.items():
        out[f'{prefix}T_{c[0]}{c[1]}_{v[0]}_{v[1]}'] = t.numpy()


def generic_case(tag, sites, v2s, lX, lY, chi, n_iter, j2=0.3, energy=True):
    """Per-function + per-move + multi-iteration comparison on a generic unit cell."""
    print(f'[{tag}]')
    dt = next(iter(sites.values())).dtype
    set_dtype(dt)
    state = IPEPS(sites, vertexToSite=v2s, lX=lX, lY=lY)
    env = ENV(chi, state)
    init_env(state, env)
    C, T = orc.init_env(sites, v2s, chi)
    for k in env.C:
        check(f'init C{k}', C[k], env.C[k], 0)
    for k in env.T:
        check(f'init T{k}', T[k], env.T[k], 0)
    out = {}
    for c, a in sites.items():
        out[f'site_{c[0]}{c[1]}'] = a.numpy()
    dump_env('init_', C, T, out)

    # one reference iteration first so that the env is not the zero-padded initial one
    args = orc.OracleArgs()
    cfg.ctm_args.ctm_max_iter = 1
    env, *_ = ctmrg.run(state, env)
    # from here on run the deterministic pieces on the REFERENCE env
    Cr = {k: v.clone() for k, v in env.C.items()}
    Tr = {k: v.clone() for k, v in env.T.items()}
    dump_env('mid_', Cr, Tr, out)
    coord = list(sites.keys())[-1]
    ref_c = dict(LU=comp.c2x2_LU, RU=comp.c2x2_RU, RD=comp.c2x2_RD, LD=comp.c2x2_LD)
    for kind in orc.CORNERS:
        r = ref_c[kind](coord, state, env, mode='sl')
        o = orc.corner_at(kind, coord, sites, v2s, Cr, Tr)
        check(f'c2x2_{kind}', o, r, 1e-13)
        out[f'c2x2_{kind}'] = r.numpy()
    ref_h = {orc.UP: comp.halves_of_4x4_CTM_MOVE_UP, orc.LEFT: comp.halves_of_4x4_CTM_MOVE_LEFT,
             orc.DOWN: comp.halves_of_4x4_CTM_MOVE_DOWN, orc.RIGHT: comp.halves_of_4x4_CTM_MOVE_RIGHT}
    ref_a = {orc.UP: ctmrg.absorb_truncate_CTM_MOVE_UP, orc.LEFT: ctmrg.absorb_truncate_CTM_MOVE_LEFT,
             orc.DOWN: ctmrg.absorb_truncate_CTM_MOVE_DOWN, orc.RIGHT: ctmrg.absorb_truncate_CTM_MOVE_RIGHT}
    for d in orc.DIRECTIONS:
        P, Pt = {}, {}
        for c in sites:
            R, Rt = ref_h[d](c, state, env, mode='sl')
            Ro, Rto = orc.halves(d, c, sites, v2s, Cr, Tr)
            check(f'halves{d}{c} R', Ro, R, 1e-13)
            check(f'halves{d}{c} Rt', Rto, Rt, 1e-13)
            P[c], Pt[c] = ctm_get_projectors_from_matrices(R, Rt, chi, cfg.ctm_args, cfg.global_args)
            Po, Pto = orc.projectors_from_matrices(R, Rt, chi, args)
            check(f'projectors{d}{c} P', Po, P[c], 1e-10)
            check(f'projectors{d}{c} Pt', Pto, Pt[c], 1e-10)
            if c == coord:
                tg = f'{d[0]}_{d[1]}'
                out[f'halves_{tg}_R'], out[f'halves_{tg}_Rt'] = R.numpy(), Rt.numpy()
                out[f'proj_{tg}_P'], out[f'proj_{tg}_Pt'] = P[c].numpy(), Pt[c].numpy()
        # absorption with the REFERENCE projectors: deterministic, compare element-wise
        for c in sites:
            r3 = ref_a[d](c, state, env, P, Pt)
            o3 = orc.absorb(d, c, sites, v2s, Cr, Tr, P, Pt, args)
            for nm, r, o in zip(('nC1', 'nC2', 'nT'), r3, o3):
                check(f'absorb{d}{c} {nm}', o, r / r.abs().max(), 1e-13)
        # the whole move, each direction from the same snapshot (SVD sign gauge may differ
        # between two evaluations of the same matrix to rounding: compare |.| as well)
        e2 = env.clone()
        ctmrg.ctm_MOVE(d, state, e2)
        C2 = {k: v.clone() for k, v in Cr.items()}
        T2 = {k: v.clone() for k, v in Tr.items()}
        orc.ctm_move(d, sites, v2s, C2, T2, chi, args)
        w_abs = max([maxrel(C2[k].abs(), e2.C[k].abs()) for k in e2.C]
                    + [maxrel(T2[k].abs(), e2.T[k].abs()) for k in e2.T])
        w_sgn = max([maxrel(C2[k], e2.C[k]) for k in e2.C] + [maxrel(T2[k], e2.T[k]) for k in e2.T])
        print(f'  move{d}: max rel diff signed {w_sgn:.2e}  |.| {w_abs:.2e}')
        assert w_abs < 1e-8
        dump_env(f'move_{d[0]}_{d[1]}_', e2.C, e2.T, out)
    for k in Cr:
        C[k] = Cr[k].clone()
    for k in Tr:
        T[k] = Tr[k].clone()
    # multi-iteration run + energy (gauge may drift: compare invariants too)
    cfg.ctm_args.ctm_max_iter = n_iter
    env, *_ = ctmrg.run(state, env)
    orc.run(sites, v2s, lX, lY, C, T, chi, n_iter, args)
    worst = 0.
    for k in env.C:
        worst = max(worst, maxrel(C[k].abs(), env.C[k].abs()))
    for k in env.T:
        worst = max(worst, maxrel(T[k].abs(), env.T[k].abs()))
    print(f'  after {n_iter} iterations max rel |C|,|T| diff   {worst:.2e}')
    assert worst < 1e-7
    dump_env('final_', env.C, env.T, out)
    spec_r = env.get_spectra()
    spec_o = orc.corner_spectra(C)
    for k in spec_r:
        check(f'spectrum C{k}', spec_o[k], spec_r[k], 1e-9)
    if energy:
        model = j1j2.J1J2(j1=1.0, j2=j2)
        check('hp', orc.j1j2_hp(1.0, j2, dt), model.get_hp((0, 0)), 1e-15)
        r = rdm.rdm2x2_legacy(coord, state, env)
        o = orc.rdm2x2(coord, sites, v2s, env.C, env.T)
        check('rdm2x2', o, r, 1e-11)
        e_ref = float(model.energy_2x2_4site(state, env))
        e_orc_on_ref = orc.energy_j1j2(sites, v2s, env.C, env.T, 1.0, j2)
        e_orc = orc.energy_j1j2(sites, v2s, C, T, 1.0, j2)
        print(f'  energy ref {e_ref:.15f} oracle(ref env) {e_orc_on_ref:.15f} oracle {e_orc:.15f}')
        assert abs(e_ref - e_orc_on_ref) < 1e-12 and abs(e_ref - e_orc) < 1e-10
        out['energy'] = np.array([e_ref])
        out['rdm2x2'] = r.numpy()
    out['meta'] = np.array(json.dumps(dict(chi=chi, n_iter=n_iter, j2=j2, lX=lX, lY=lY)))
    np.savez_compressed(os.path.join(GOLD, f'{tag}.npz'), **out)


def c4v_case(tag, a, chi, n_iter, j2=0.3):
    print(f'[{tag}]')
    set_dtype(a.dtype)
    state = IPEPS_C4V(a)
    env = ENV_C4V(chi, state)
    env_c4v.init_env(state, env)
    C, T = orc.init_env_c4v(a, chi)
    check('init C', C, env.C[env.keyC], 1e-14)
    check('init |T|', T.abs(), env.T[env.keyT].abs(), 1e-12)
    out = dict(site=a.numpy(), init_C=env.C[env.keyC].numpy(), init_T=env.T[env.keyT].numpy())
    C, T = env.C[env.keyC].clone(), env.T[env.keyT].clone()   # identical start (eigh gauge)

    def f_eig(M, chi_):
        return truncated_eig_sym(M, chi_, keep_multiplets=True)
    for i in range(n_iter):
        if i == 2:
            Cm, Tm = env.C[env.keyC].clone(), env.T[env.keyT].clone()
            r = c2x2_sl(a, Cm, Tm)
            check('c2x2_sl', orc.c2x2_c4v(a, Cm, Tm), r, 1e-13)
            out['mid_C'], out['mid_T'], out['mid_c2x2'] = Cm.numpy(), Tm.numpy(), r.numpy()
        ctmrg_c4v.ctm_MOVE_sl(a, env, f_eig)
        if i == 2:
            nC, nT = orc.ctm_move_c4v(a, Cm, Tm, chi)
            check('single move C', nC, env.C[env.keyC], 1e-12)
            check('single move |T|', nT.abs(), env.T[env.keyT].abs(), 1e-10)
            out['mid_nC'], out['mid_nT'] = env.C[env.keyC].numpy(), env.T[env.keyT].numpy()
        C, T = orc.ctm_move_c4v(a, C, T, chi)
    check(f'C after {n_iter} moves', C, env.C[env.keyC], 1e-10)
    check(f'|T| after {n_iter} moves', T.abs(), env.T[env.keyT].abs(), 1e-9)
    out['final_C'], out['final_T'] = env.C[env.keyC].numpy(), env.T[env.keyT].numpy()
    model = j1j2.J1J2_C4V_BIPARTITE(j1=1.0, j2=j2)
    e_ref = float(model.energy_1x1_lowmem(state, env))
    e_orc = orc.energy_j1j2_c4v(a, env.C[env.keyC], env.T[env.keyT], 1.0, j2)
    print(f'  energy ref {e_ref:.15f} oracle(ref env) {e_orc:.15f}')
    assert abs(e_ref - e_orc) < 1e-11
    out['energy'] = np.array([e_ref])
    out['meta'] = np.array(json.dumps(dict(chi=chi, n_iter=n_iter, j2=j2)))
    np.savez_compressed(os.path.join(GOLD, f'{tag}.npz'), **out)


def rvb_known_answer():
    """TestRVB.test_ctmrg_RVB (examples/j1j2/ctmrg_j1j2_c4v.py:214-259): RVB_1x1.in, D=3,
    chi=16, j2=0.5 -> e = -0.47684229 @1e-8.  Run through the ORACLE; store the state."""
    print('[rvb known answer]')
    from ipeps.ipeps_c4v import read_ipeps_c4v
    set_dtype(torch.float64)
    state = read_ipeps_c4v(os.path.join(REF, 'test-input', 'RVB_1x1.in'))
    a = state.site()
    chi = 16
    C, T = orc.init_env_c4v(a, chi)
    e_prev = 0
    for i in range(200):
        C, T = orc.ctm_move_c4v(a, C, T, chi)
        e = orc.energy_j1j2_c4v(a, C, T, 1.0, 0.5)
        if abs(e - e_prev) < 1e-11:
            break
        e_prev = e
    print(f'  oracle energy {e:.10f} after {i + 1} moves (reference test value -0.47684229)')
    assert abs(e - (-0.47684229)) < 1e-8
    np.savez_compressed(os.path.join(GOLD, 'rvb_c4v_known_answer.npz'), site=a.numpy(),
                        energy=np.array([-0.47684229]), chi=np.array([chi]), j2=np.array([0.5]))


def states_known_answer():
    """TestCtmrg_States, 2SITE ansatz (examples/j1j2/ctmrg_j1j2.py:259-267): state
    gesdd-D2-chi50-j20.55-run0-iRND2x1_state.json, chi=32, j2=0.55, energy-converged CTM
    (ctm_conv_tol 1e-8, <=100 iterations): FINAL energy -0.4434603770143078 @1e-6.
    Run through the ORACLE (generic move + rdm2x2 energy); the state is stored as npz."""
    print('[TestCtmrg_States 2SITE known answer]')
    from ipeps.ipeps import read_ipeps
    from collections import OrderedDict
    set_dtype(torch.float64)
    state = read_ipeps(os.path.join(REF, 'test-input', 'gesdd-D2-chi50-j20.55-run0-iRND2x1_state.json'),
                       vertexToSite=orc.v2s_2site)
    sites = OrderedDict((c, t.clone()) for c, t in state.sites.items())
    chi, j2 = 32, 0.55
    C, T = orc.init_env(sites, orc.v2s_2site, chi)
    e_prev = None
    for i in range(100):
        orc.ctm_iteration(sites, orc.v2s_2site, state.lX, state.lY, C, T, chi)
        e = orc.energy_j1j2(sites, orc.v2s_2site, C, T, 1.0, j2)
        if e_prev is not None and abs(e - e_prev) < 1e-8:
            break
        e_prev = e
    print(f'  oracle energy {e:.12f} after {i + 1} iterations (reference test value -0.4434603770143078)')
    assert abs(e - (-0.4434603770143078)) < 1e-6
    out = {f'site_{c[0]}{c[1]}': t.numpy() for c, t in sites.items()}
    np.savez_compressed(os.path.join(GOLD, 'j1j2_2site_known_answer.npz'), energy=np.array([-0.4434603770143078]),
                        chi=np.array([chi]), j2=np.array([j2]), lX=np.array([state.lX]), lY=np.array([state.lY]), **out)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    generic_case('generic_4site_D2_chi8_A', orc.random_state_4site(2, family='A'), orc.v2s_4site, 2, 2, 8, 3)
    generic_case('generic_4site_D2_chi8_B', orc.random_state_4site(2, family='B'), orc.v2s_4site, 2, 2, 8, 3)
    generic_case('generic_4site_D3_chi12_B', orc.random_state_4site(3, family='B'), orc.v2s_4site, 2, 2, 12, 2)
    from collections import OrderedDict
    generic_case('kagome_1site_D2_chi8_A', OrderedDict({(0, 0): orc.random_state_kagome(2, family='A')}),
                 orc.v2s_1site, 1, 1, 8, 3, energy=False)
    sB = orc.random_state_4site(2, family='B', dtype=torch.complex128)
    generic_case('generic_4site_D2_chi8_B_c128', sB, orc.v2s_4site, 2, 2, 8, 2)
    c4v_case('c4v_D2_chi8_A', orc.random_state_c4v(2, family='A'), 8, 6)
    c4v_case('c4v_D2_chi8_B', orc.random_state_c4v(2, family='B'), 8, 6)
    c4v_case('c4v_D2_chi8_B_c128', orc.random_state_c4v(2, family='B', dtype=torch.complex128), 8, 6)
    rvb_known_answer()
    states_known_answer()
    print('all reference-vs-oracle checks passed; fixtures in', os.path.abspath(GOLD))
