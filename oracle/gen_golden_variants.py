"""Reference outputs for the variants of the move and for the density matrices -> tests/golden/variants_*.npz.

Run in the BUILD container only (needs /root/reference, which never travels):

    python oracle/gen_golden_variants.py

Inputs are the states and mid-run environments of the fixtures written by oracle/gen_golden.py.  Every array stored here
is an output of the UNMODIFIED reference (ctm.generic.ctmrg.ctm_MOVE under projector_method='4X2', under ctm_force_dl
with rank-4 sites, under ctm_absorb_normalization='fro'; ctm.generic.rdm.rdm2x2_legacy / rdm1x1_dl / rdm2x1_dl / rdm1x2_dl;
ctm.one_site_c4v.ctmrg_c4v.ctm_MOVE_dl; ctm.one_site_c4v.rdm_c4v.*).  The oracle is compared with them on the way (the
script aborts on a mismatch); tests/ re-check the oracle (CPU) and libctmb (GPU box) against the stored arrays.
"""
import copy
import os
import sys
import json
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
GOLD = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
os.chdir('/tmp')  # config.configure writes log files into cwd

import ctm_oracle as orc                                     # noqa: E402
import helpers as H                                           # noqa: E402
import config as cfg                                          # noqa: E402
from ipeps.ipeps import IPEPS                                 # noqa: E402
from ipeps.ipeps_c4v import IPEPS_C4V                         # noqa: E402
from ctm.generic.env import ENV                               # noqa: E402
from ctm.generic import ctmrg, rdm                            # noqa: E402
from ctm.one_site_c4v.env_c4v import ENV_C4V                  # noqa: E402
from ctm.one_site_c4v import ctmrg_c4v, rdm_c4v               # noqa: E402
from linalg.custom_eig import truncated_eig_sym               # noqa: E402

DEFAULTS = copy.deepcopy(cfg.ctm_args.__dict__)


def reset(dtype):
    cfg.ctm_args.__dict__.clear()
    cfg.ctm_args.__dict__.update(copy.deepcopy(DEFAULTS))
    cfg.global_args.dtype = 'complex128' if dtype.is_complex else 'float64'
    cfg.global_args.torch_dtype = dtype
    cfg.global_args.device = 'cpu'


def worst_abs(C1, T1, C2, T2, keys):
    return max(float((X[k].abs() - Y[k].abs()).abs().max()) for X, Y in ((C1, C2), (T1, T2)) for k in keys if k in Y)


def key_str(k):
    (c, v) = k
    return f'{c[0]}{c[1]}_{v[0]}_{v[1]}'


def generic(name):
    print(f'[{name}]')
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'mid_')
    dt = sites[(0, 0)].dtype
    dl = type(sites)((c, orc.double_layer(a)) for c, a in sites.items())
    out = {'meta': json.dumps({'source': name, 'chi': chi, 'env': 'mid_'})}
    variants = {'4x2': (dict(projector_method='4X2'), sites, orc.OracleArgs(projector_method='4X2')),
                'dl': (dict(ctm_force_dl=True), dl, orc.OracleArgs()),
                'fro': (dict(ctm_absorb_normalization='fro'), sites, orc.OracleArgs(ctm_absorb_normalization='fro'))}
    for tag, (ref_kw, ref_sites, oargs) in variants.items():
        reset(dt)
        for k, v in ref_kw.items():
            setattr(cfg.ctm_args, k, v)
        state = IPEPS(sites={c: t.clone() for c, t in ref_sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
        for d in orc.DIRECTIONS:
            env = ENV(chi, state)
            env.C = {k: v.clone() for k, v in C0.items()}
            env.T = {k: v.clone() for k, v in T0.items()}
            ctmrg.ctm_MOVE(d, state, env, ctm_args=cfg.ctm_args, global_args=cfg.global_args)
            C, T = dict(C0), dict(T0)
            orc.ctm_move(d, ref_sites, v2s, C, T, chi, oargs)
            kC1, kC2, kT = orc.ABSORB[d]['out']
            changed = [k for k in list(C0) + list(T0) if k[1] in (kC1, kC2, kT)]
            err = worst_abs(C, T, env.C, env.T, changed)
            print(f'  move {tag:4s} {str(d):8s} oracle vs reference |.| {err:.2e}')
            assert err < 1e-9
            for k in changed:
                t = env.C[k] if k in env.C and k[1] in (kC1, kC2) else env.T[k]
                out[f'move_{tag}_{d[0]}_{d[1]}_{"C" if t.dim() == 2 else "T"}_{key_str(k)}'] = t.numpy()
    reset(dt)
    state = IPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    env = ENV(chi, state)
    env.C, env.T = dict(C0), dict(T0)
    for coord in sites:
        for spd in (False, True):
            for fname, f_ref, f_orc in (('rdm2x2', rdm.rdm2x2_legacy, orc.rdm2x2), ('rdm1x1', rdm.rdm1x1_dl, orc.rdm1x1),
                                        ('rdm2x1', rdm.rdm2x1_dl, orc.rdm2x1), ('rdm1x2', rdm.rdm1x2_dl, orc.rdm1x2)):
                r = f_ref(coord, state, env, sym_pos_def=spd)
                o = f_orc(coord, sites, v2s, C0, T0, sym_pos_def=spd)
                assert float((r - o).abs().max()) < 1e-13, (fname, coord, spd)
                out[f'{fname}_{coord[0]}{coord[1]}_{int(spd)}'] = r.numpy()
    print('  density matrices: oracle vs reference < 1e-13')
    np.savez_compressed(os.path.join(GOLD, f'variants_{name}.npz'), **out)


def c4v(name):
    print(f'[{name}]')
    z, meta = H.load_golden(name)
    chi = meta['chi']
    a = torch.from_numpy(z['site'])
    reset(a.dtype)
    out = {'meta': json.dumps({'source': name, 'chi': chi, 'n_moves_dl': 3})}
    A = orc.double_layer(a)
    C, T = torch.from_numpy(z['init_C']), torch.from_numpy(z['init_T'])
    state = IPEPS_C4V(a.clone())
    env = ENV_C4V(chi, state)
    env.C[env.keyC], env.T[env.keyT] = C.clone(), T.clone()
    for _ in range(3):
        ctmrg_c4v.ctm_MOVE_dl(A, env, lambda M, chi_: truncated_eig_sym(M, chi_, keep_multiplets=True),
                              ctm_args=cfg.ctm_args, global_args=cfg.global_args)
        C, T = orc.ctm_move_c4v(A, C, T, chi)
    assert float((env.get_C() - C).abs().max()) < 1e-12 and float((env.get_T().abs() - T.abs()).abs().max()) < 1e-10
    out['dl3_C'], out['dl3_T'] = env.get_C().numpy(), env.get_T().numpy()
    Cf, Tf = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    env.C[env.keyC], env.T[env.keyT] = Cf.clone(), Tf.clone()
    for spd in (False, True):
        for fname, f_ref, o in (('rdm2x2_NN', rdm_c4v.rdm2x2_NN_lowmem_sl, orc.rdm2x2_c4v(a, Cf, Tf, (0, 1), spd)),
                                ('rdm2x2_NNN', rdm_c4v.rdm2x2_NNN_lowmem_sl, orc.rdm2x2_c4v(a, Cf, Tf, (0, 3), spd)),
                                ('rdm2x2', rdm_c4v.rdm2x2, orc.rdm2x2_c4v(a, Cf, Tf, (0, 1, 2, 3), spd)),
                                ('rdm1x1', rdm_c4v.rdm1x1_sl, orc.rdm_small_c4v('1x1', a, Cf, Tf, spd)),
                                ('rdm2x1', rdm_c4v.rdm2x1_sl, orc.rdm_small_c4v('2x1', a, Cf, Tf, spd))):
            r = f_ref(state, env, sym_pos_def=spd)
            assert float((r - o).abs().max()) < 1e-12, (fname, spd)
            out[f'{fname}_{int(spd)}'] = r.numpy()
    print('  ctm_MOVE_dl x3 and density matrices: oracle vs reference ok')
    np.savez_compressed(os.path.join(GOLD, f'variants_{name}.npz'), **out)


if __name__ == '__main__':
    for n in ('generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128'):
        generic(n)
    for n in ('c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'):
        c4v(n)
    print('variant fixtures written to', os.path.abspath(GOLD))
