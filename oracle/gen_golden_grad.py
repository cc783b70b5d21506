"""Gradients written by the UNMODIFIED reference -> tests/golden/grad_*.npz  (AD through the move).

Run in the BUILD container only (needs /root/reference, which never travels):

    python oracle/gen_golden_grad.py

For each case: a state, a fixed (detached) starting environment, N CTM moves of the reference recorded by autograd
(ctm.one_site_c4v.ctmrg_c4v.run / ctm.generic.ctmrg.run with conv_check=None: SYMEIG.backward linalg/eig_sym.py:56-78,
SVDGESDD.backward linalg/svd_gesdd.py:209-328), the J1-J2 energy, loss.backward().  Two losses are differentiated:
the reference model's own energy (models/j1j2.py: J1J2_C4V_BIPARTITE.energy_1x1_lowmem, J1J2.energy_2x2_4site through
rdm2x2_legacy) and the oracle's restatement of it (ctm_oracle.energy_j1j2*, as_tensor=True) on the reference's
environment; the script aborts unless the two gradients agree, and stores the reference's.  tests/ then differentiate
the oracle's energy on the environment produced by peps_torch_b200/ad.py -- everything between the state and the
environment is the code under test, everything after it is identical in fixture and test.
"""
import copy
import os
import sys
import json
from collections import OrderedDict
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
GOLD = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
os.chdir('/tmp')

import ctm_oracle as orc                                     # noqa: E402
import config as cfg                                          # noqa: E402
from ipeps.ipeps import IPEPS                                 # noqa: E402
from ipeps.ipeps_c4v import IPEPS_C4V                         # noqa: E402
from ctm.generic.env import ENV                               # noqa: E402
from ctm.generic import ctmrg, rdm                            # noqa: E402
from ctm.one_site_c4v.env_c4v import ENV_C4V                  # noqa: E402
from ctm.one_site_c4v import ctmrg_c4v                        # noqa: E402
from models import j1j2                                       # noqa: E402

rdm.rdm2x2 = rdm.rdm2x2_legacy            # opt_einsum is absent (SURVEY 8c caveat 1)
DEFAULTS = copy.deepcopy(cfg.ctm_args.__dict__)


def reset(dtype):
    cfg.ctm_args.__dict__.clear()
    cfg.ctm_args.__dict__.update(copy.deepcopy(DEFAULTS))
    cfg.global_args.dtype = 'complex128' if dtype.is_complex else 'float64'
    cfg.global_args.torch_dtype = dtype
    cfg.global_args.device = 'cpu'


def symm(A):
    """ipeps/ipeps_c4v.py:90-91 / groups/pg.py:44-69 through the oracle's projectors (differentiable torch ops)."""
    if A.is_complex():
        return orc.make_c4v_symm_A1(A.real) + 1j * orc.make_c4v_symm_A2(A.imag)
    return orc.make_c4v_symm_A1(A)


def c4v_case(name, D, chi, dtype, pre, moves, j2, checkpoint=False):
    reset(dtype)
    a0 = orc.random_state_c4v(D, family='B', dtype=dtype)
    C0, T0 = orc.init_env_c4v(a0, chi)
    for _ in range(pre):                                     # a generic, full-rank starting environment
        C0, T0 = orc.ctm_move_c4v(a0, C0, T0, chi)
    model = j1j2.J1J2_C4V_BIPARTITE(j1=1.0, j2=j2)
    cfg.ctm_args.ctm_max_iter = moves
    cfg.ctm_args.fwd_checkpoint_move = checkpoint
    grads = []
    for which in ('reference', 'oracle'):
        # the parameter is projected on the C4v-symmetric subspace inside the graph, as optim_j1j2_c4v.py does with
        # to_ipeps_c4v (the reference's C4v energy rotates ONE enlarged corner: as a function of an unsymmetric tensor it
        # is a different function, so only the projected gradient is defined by the model)
        A = a0.clone().requires_grad_(True)
        a = symm(A)
        state = IPEPS_C4V(a)
        env = ENV_C4V(chi, state)
        env.C[env.keyC], env.T[env.keyT] = C0.clone(), T0.clone()
        env, *_ = ctmrg_c4v.run(state, env, conv_check=None)
        if which == 'reference':
            loss = model.energy_1x1_lowmem(state, env)
        else:
            loss = orc.energy_j1j2_c4v(a, env.C[env.keyC], env.T[env.keyT], 1.0, j2, as_tensor=True)
        loss.backward()
        grads.append((float(loss.detach().real), A.grad.clone()))
    (e_ref, g_ref), (e_orc, g_orc) = grads
    scale = float(g_ref.abs().max())
    d = float((g_ref - g_orc).abs().max())
    print(f'[{name}] energy {e_ref:.15f} (oracle energy {e_orc:.15f})  |grad| {scale:.3e}  oracle-energy grad vs reference-energy grad {d:.2e}')
    assert abs(e_ref - e_orc) < 1e-12 and d < 1e-10 * max(1.0, scale), (e_ref, e_orc, d)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), site=a0.numpy(), C0=C0.numpy(), T0=T0.numpy(), grad=g_ref.numpy(),
                        energy=np.array([e_ref]),
                        meta=json.dumps(dict(kind='c4v', D=D, chi=chi, moves=moves, j2=j2, pre=pre,
                                             ad_decomp_reg=cfg.ctm_args.ad_decomp_reg)))


def generic_case(name, D, chi, dtype, moves_iter, j2):
    reset(dtype)
    sites0 = orc.random_state_4site(D, family='B', dtype=dtype)
    C0, T0 = orc.init_env(sites0, orc.v2s_4site, chi)
    orc.ctm_iteration(sites0, orc.v2s_4site, 2, 2, C0, T0, chi)
    model = j1j2.J1J2(j1=1.0, j2=j2)
    cfg.ctm_args.ctm_max_iter = moves_iter
    grads = []
    for which in ('reference', 'oracle'):
        sites = OrderedDict((c, t.clone().requires_grad_(True)) for c, t in sites0.items())
        state = IPEPS(sites, vertexToSite=orc.v2s_4site, lX=2, lY=2)
        env = ENV(chi, state)
        env.C = {k: v.clone() for k, v in C0.items()}
        env.T = {k: v.clone() for k, v in T0.items()}
        env, *_ = ctmrg.run(state, env, conv_check=None)
        if which == 'reference':
            loss = model.energy_2x2_4site(state, env)
        else:
            loss = orc.energy_j1j2(sites, orc.v2s_4site, env.C, env.T, 1.0, j2, as_tensor=True)
        loss.backward()
        grads.append((float(loss.detach().real), {c: t.grad.clone() for c, t in sites.items()}))
    (e_ref, g_ref), (e_orc, g_orc) = grads
    scale = max(float(g.abs().max()) for g in g_ref.values())
    d = max(float((g_ref[c] - g_orc[c]).abs().max()) for c in g_ref)
    print(f'[{name}] energy {e_ref:.15f} (oracle energy {e_orc:.15f})  |grad| {scale:.3e}  oracle-energy grad vs reference-energy grad {d:.2e}')
    assert abs(e_ref - e_orc) < 1e-12 and d < 1e-10 * max(1.0, scale), (e_ref, e_orc, d)
    out = dict(energy=np.array([e_ref]),
               meta=json.dumps(dict(kind='generic', D=D, chi=chi, iters=moves_iter, j2=j2, ad_decomp_reg=cfg.ctm_args.ad_decomp_reg)))
    for c, t in sites0.items():
        out[f'site_{c[0]}{c[1]}'] = t.numpy()
        out[f'grad_{c[0]}{c[1]}'] = g_ref[c].numpy()
    for k, v in C0.items():
        out[f'C0_{k[0][0]}{k[0][1]}_{k[1][0]}_{k[1][1]}'] = v.numpy()
    for k, v in T0.items():
        out[f'T0_{k[0][0]}{k[0][1]}_{k[1][0]}_{k[1][1]}'] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)


if __name__ == '__main__':
    c4v_case('grad_c4v_D2_chi16', 2, 16, torch.float64, pre=3, moves=4, j2=0.3)
    c4v_case('grad_c4v_D2_chi16_c128', 2, 16, torch.complex128, pre=3, moves=3, j2=0.3)
    c4v_case('grad_c4v_D2_chi16_ckpt', 2, 16, torch.float64, pre=3, moves=2, j2=0.3, checkpoint=True)
    generic_case('grad_generic_4site_D2_chi8', 2, 8, torch.float64, 1, 0.3)
    generic_case('grad_generic_4site_D2_chi6_c128', 2, 6, torch.complex128, 1, 0.3)
