"""Host-side containers and convergence criterion of peps_torch_b200/env.py against the unmodified reference
(ctm/generic/env.py, ctm/one_site_c4v/env_c4v.py).  Needs the reference tree (build container only)."""
import copy
import os
import sys
import pytest
import torch
import ctm_oracle as orc
import helpers as H

REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ctm', 'generic')), reason='reference tree not present')


@pytest.fixture()
def ref(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import config as cfg
    saved = copy.deepcopy(cfg.ctm_args.__dict__)
    cfg.global_args.dtype, cfg.global_args.device = 'float64', 'cpu'
    try:
        yield cfg
    finally:
        cfg.ctm_args.__dict__.clear()
        cfg.ctm_args.__dict__.update(saved)
        os.chdir(cwd)
        sys.path.remove(REF)


def _generic(name='generic_4site_D2_chi8_B'):
    z, meta = H.load_golden(name)
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    return sites, v2s, lX, lY, meta['chi'], z


def test_env_methods_match_reference(ref):
    from ipeps.ipeps import IPEPS as RefIPEPS
    from ctm.generic.env import ENV as RefENV, init_env as ref_init_env
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    sites, v2s, lX, lY, chi, z = _generic()
    rs = RefIPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    renv = RefENV(chi, rs)
    ref_init_env(rs, renv)
    st = IPEPS(sites, v2s, lX, lY)
    env = ENV(chi, st)
    init_env(st, env)
    assert set(env.C) == set(renv.C) and set(env.T) == set(renv.T)
    assert all(torch.equal(env.C[k], renv.C[k]) for k in renv.C) and all(torch.equal(env.T[k], renv.T[k]) for k in renv.T)
    C, T = H.golden_env(z, 'mid_')
    env.C, env.T, renv.C, renv.T = dict(C), dict(T), dict(C), dict(T)
    for new_chi in (chi + 5, chi - 3, chi):
        a, b = env.extend(new_chi), renv.extend(new_chi)
        assert a.chi == b.chi == new_chi
        assert all(torch.equal(a.C[k], b.C[k]) for k in b.C) and all(torch.equal(a.T[k], b.T[k]) for k in b.T)
    assert env.min_chi() == renv.min_chi() == chi
    assert env.extend(chi - 3).min_chi() == chi - 3
    for coord in ((0, 0), (3, -2)):
        for x, y in zip(env.get_site_env_t(coord, st), renv.get_site_env_t(coord, rs)):
            assert x is y or torch.equal(x, y)
    sa, sb = env.get_spectra(), renv.get_spectra()
    assert all(torch.allclose(sa[k], sb[k], atol=1e-15) for k in sb)
    c = env.clone()
    assert all(torch.equal(c.C[k], env.C[k]) and c.C[k] is not env.C[k] for k in env.C)
    d = env.detach()
    assert all(d.T[k].data_ptr() == env.T[k].data_ptr() for k in env.T)
    # RANDOM initialisation: same shapes / dtype, values in [0,1)
    ref.ctm_args.ctm_env_init_type = 'RANDOM'
    from peps_torch_b200.config import CTMARGS
    args = CTMARGS(); args.ctm_env_init_type = 'RANDOM'
    e2 = ENV(chi, st)
    init_env(st, e2, ctm_args=args)
    assert all(e2.T[k].shape == renv.T[k].shape and 0 <= float(e2.T[k].min()) and float(e2.T[k].max()) < 1 for k in renv.T)
    args.ctm_env_init_type = 'NOPE'
    with pytest.raises(ValueError):
        init_env(st, e2, ctm_args=args)
    dl = IPEPS({c_: orc.double_layer(t) for c_, t in sites.items()}, v2s, lX, lY)
    args.ctm_env_init_type = 'CTMRG'
    with pytest.raises(RuntimeError):
        init_env(dl, e2, ctm_args=args)


def test_conv_specC_matches_reference_over_a_run(ref):
    """Same history and the same converged / not-converged decisions as env.py:816-875 along a CTM run."""
    from ipeps.ipeps import IPEPS as RefIPEPS
    from ctm.generic.env import ENV as RefENV, ctmrg_conv_specC as ref_conv
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, ctmrg_conv_specC
    from peps_torch_b200.config import CTMARGS
    sites, v2s, lX, lY, chi, z = _generic('generic_4site_D2_chi8_A')
    C, T = H.golden_env(z, 'init_')
    rs = RefIPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    st = IPEPS(sites, v2s, lX, lY)
    for p in ('inf', 'fro'):
        Cw, Tw = dict(C), dict(T)
        renv, env = RefENV(chi, rs), ENV(chi, st)
        args = CTMARGS(); args.ctm_conv_tol = 1e-6; args.ctm_max_iter = 12
        ref.ctm_args.ctm_conv_tol, ref.ctm_args.ctm_max_iter = 1e-6, 12
        h_ref, h = None, None
        for it in range(12):
            orc.ctm_iteration(sites, v2s, lX, lY, Cw, Tw, chi)
            renv.C, renv.T, env.C, env.T = dict(Cw), dict(Tw), dict(Cw), dict(Tw)
            done_ref, h_ref = ref_conv(rs, renv, h_ref, p=p, ctm_args=ref.ctm_args)
            done, h = ctmrg_conv_specC(st, env, h, p=p, ctm_args=args)
            assert done == done_ref, (p, it)
            a, b = h['conv_crit'][-1], h_ref['conv_crit'][-1]
            assert a == b or abs(a - b) <= 1e-12 * max(abs(b), 1e-300) + 1e-18, (p, it, a, b)
            if done:
                break
        assert done and len(h['conv_crit']) == len(h_ref['conv_crit']) < 12      # converged before the iteration cap


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'rvb_c4v_known_answer'])
def test_env_c4v_methods_match_reference(ref, name):
    from ctm.one_site_c4v.env_c4v import ENV_C4V as RefENV, compute_multiplets as ref_mult, init_env as ref_init
    from ipeps.ipeps_c4v import IPEPS_C4V as RefState
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V, init_env_c4v, compute_multiplets
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    chi = meta['chi'] if 'chi' in meta else int(z['chi'])
    if 'final_C' not in z.files:             # known-answer fixture: state only -> converge a few moves with the oracle
        Cf, Tf = orc.init_env_c4v(a, chi)
        for _ in range(6):
            Cf, Tf = orc.ctm_move_c4v(a, Cf, Tf, chi)
    else:
        Cf, Tf = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    rs, st = RefState(a.clone()), IPEPS_C4V(a)
    renv, env = RefENV(chi, rs), ENV_C4V(chi, st)
    ref_init(rs, renv)
    init_env_c4v(st, env)
    assert H.maxrel(env.get_C(), renv.get_C()) < 1e-13 and H.maxrel(env.get_T().abs(), renv.get_T().abs()) < 1e-12
    for e in (renv, env):
        e.C[e.keyC], e.T[e.keyT] = Cf.clone(), Tf.clone()
    assert compute_multiplets(env) == ref_mult(renv)
    for new_chi in (chi + 4, chi - 2):
        x, y = env.extend(new_chi), renv.extend(new_chi)
        assert torch.equal(x.get_C(), y.get_C()) and torch.equal(x.get_T(), y.get_T()) and x.chi == new_chi
    c = env.clone()
    assert torch.equal(c.get_C(), env.get_C()) and c.get_C() is not env.get_C() and c.bond_dim == env.bond_dim
    # custom (C, T) smaller than chi: zero-padded
    e3, r3 = ENV_C4V(chi + 3, st), RefENV(chi + 3, rs)
    init_env_c4v(st, e3, C_and_T=(Cf, Tf))
    ref_init(rs, r3, C_and_T=(Cf, Tf))
    assert torch.equal(e3.get_C(), r3.get_C()) and torch.equal(e3.get_T(), r3.get_T())


@pytest.mark.parametrize('kind', ['PROD', 'CTMRG_OBC', 'RANDOM'])
@pytest.mark.parametrize('dl', [False, True])
def test_env_initialisations_match_reference(ref, kind, dl):
    """ctm_env_init_type 'PROD' (env.py:274-365) and 'CTMRG_OBC' (:538-715) element for element, on single-layer and on
    double-layer (rank-4) sites; 'RANDOM' through shapes and dtype (the draws differ)."""
    from ipeps.ipeps import IPEPS as RefIPEPS
    from ctm.generic.env import ENV as RefENV, init_env as ref_init_env
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    from peps_torch_b200.config import CTMARGS
    for name in ('generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128'):
        sites, v2s, lX, lY, chi, z = _generic(name)
        if dl:
            sites = type(sites)((c, orc.double_layer(a)) for c, a in sites.items())
        ref.global_args.dtype = 'complex128' if next(iter(sites.values())).is_complex() else 'float64'
        ref.global_args.torch_dtype = next(iter(sites.values())).dtype
        for chi_ in (chi, 3):                       # 3 < D^2: the truncating branch of the OBC copy
            rs = RefIPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
            renv = RefENV(chi_, rs)
            ref.ctm_args.ctm_env_init_type = kind
            ref_init_env(rs, renv, ctm_args=ref.ctm_args)
            st = IPEPS(sites, v2s, lX, lY)
            env = ENV(chi_, st)
            args = CTMARGS(); args.ctm_env_init_type = kind
            init_env(st, env, args)
            assert set(env.C) == set(renv.C) and set(env.T) == set(renv.T)
            for a, b in [(env.C[k], renv.C[k]) for k in renv.C] + [(env.T[k], renv.T[k]) for k in renv.T]:
                assert a.shape == b.shape and a.dtype == b.dtype
                if kind != 'RANDOM':
                    assert float((a - b).abs().max()) < 1e-15, (kind, dl, name, chi_)
    ref.global_args.dtype, ref.global_args.torch_dtype = 'float64', torch.float64
