"""Transfer-operator mat-vecs and leading spectra: peps_torch_b200/ctm/generic/{corrf,transferops}.py against
the unmodified reference (ctm/generic/corrf.py:278-650, transferops.py:14-205) with the oracle standing in for libctmb.
Needs the reference tree (build container only); the GPU suite compares libctmb with the oracle-engine result."""
import copy
import os
import sys
import pytest
import torch
import helpers as H

REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ctm', 'generic')), reason='reference tree not present')
DIRS = [(0, -1), (-1, 0), (0, 1), (1, 0)]


@pytest.fixture()
def ref(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import config as cfg
    saved = copy.deepcopy(cfg.global_args.__dict__)
    try:
        yield cfg
    finally:
        cfg.global_args.__dict__.clear()
        cfg.global_args.__dict__.update(saved)
        os.chdir(cwd)
        sys.path.remove(REF)


def edge_shapes(st, env, d):
    from peps_torch_b200.ctm.generic import transferops as ot
    chi1, chi2 = ot._get_chis(st, env, (0, 0), d, 0)
    leg = {(0, -1): 1, (-1, 0): 2, (0, 1): 3, (1, 0): 4}[(-d[0], -d[1])]
    return chi1, st.sites[(0, 0)].shape[leg] ** 2, chi2


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'])
def test_transfer_operators_match_reference(ref, name, monkeypatch):
    from ipeps.ipeps import IPEPS as RI
    from ctm.generic.env import ENV as RE
    from ctm.generic import transferops as rt, corrf as rc
    from peps_torch_b200.ctm.generic import transferops as ot, corrf as oc
    eng = H.OracleEngine()
    monkeypatch.setattr(oc, '_engine', lambda: eng)
    z, meta = H.load_golden(name)
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'final_' if any(k.startswith('final_') for k in z.files) else 'mid_')
    dt = next(iter(sites.values())).dtype
    ref.global_args.dtype, ref.global_args.torch_dtype, ref.global_args.device = ('complex128' if dt.is_complex else 'float64'), dt, 'cpu'
    rs = RI(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    re = RE(meta['chi'], rs)
    re.C, re.T = dict(C), dict(T)
    st, env = H.State(sites, v2s, lX, lY), H.Env(meta['chi'], dict(C), dict(T))
    for d in DIRS:
        g = torch.Generator().manual_seed(1)
        chi1, d2, chi2 = edge_shapes(st, env, d)
        V = torch.randn(chi1, d2, chi2, dtype=dt, generator=g)
        assert H.maxrel(oc.apply_TM_1sO((0, 0), d, st, env, V), rc.apply_TM_1sO((0, 0), d, rs, re, V)) < 1e-13
        V0 = torch.randn(chi1, chi2, dtype=dt, generator=g)
        assert H.maxrel(oc.apply_TM_0sO((0, 0), d, st, env, V0), rc.apply_TM_0sO((0, 0), d, rs, re, V0)) < 1e-13
        La, Lb = ot.get_Top_spec(3, (0, 0), d, st, env), rt.get_Top_spec(3, (0, 0), d, rs, re)
        assert float((La - Lb).abs().max()) < 1e-10, (d, La, Lb)
        Wa, Wb = ot.get_Top_w0_spec(3, (0, 0), d, st, env), rt.get_Top_w0_spec(3, (0, 0), d, rs, re)
        assert float((Wa.abs() - Wb.abs()).abs().max()) < 1e-10, (d, Wa, Wb)


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'])
def test_edges_operator_insertions_and_corrf_match_reference(ref, name, monkeypatch):
    """get_edge, apply_edge, apply_TM_1sO with a one-site operator and corrf_1sO1sO (corrf.py:10-104, 234-277, 415-419,
    980-1067), with and without the rl_0 eigenvector edges, in all four directions."""
    from ipeps.ipeps import IPEPS as RI
    from ctm.generic.env import ENV as RE
    from ctm.generic import corrf as rc
    from peps_torch_b200.ctm.generic import corrf as oc
    eng = H.OracleEngine()
    monkeypatch.setattr(oc, '_engine', lambda: eng)
    z, meta = H.load_golden(name)
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'final_' if any(k.startswith('final_') for k in z.files) else 'mid_')
    dt = next(iter(sites.values())).dtype
    ref.global_args.dtype, ref.global_args.torch_dtype, ref.global_args.device = ('complex128' if dt.is_complex else 'float64'), dt, 'cpu'
    rs = RI(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    re = RE(meta['chi'], rs)
    re.C, re.T = dict(C), dict(T)
    st, env = H.State(sites, v2s, lX, lY), H.Env(meta['chi'], dict(C), dict(T))
    p = next(iter(sites.values())).shape[0]
    g = torch.Generator().manual_seed(5)
    op1 = torch.randn(p, p, dtype=dt, generator=g)
    ops2 = [torch.randn(p, p, dtype=dt, generator=g) for _ in range(4)]
    for d in DIRS:
        rev = (-d[0], -d[1])
        assert H.maxrel(oc.get_edge((0, 0), d, st, env), rc.get_edge((0, 0), d, rs, re)) < 1e-13
        chi1, d2, chi2 = edge_shapes(st, env, d)
        V = torch.randn(chi1, d2, chi2, dtype=dt, generator=g)
        assert H.maxrel(oc.apply_TM_1sO((0, 0), d, st, env, V, op=op1), rc.apply_TM_1sO((0, 0), d, rs, re, V, op=op1)) < 1e-13
        Ve = torch.randn(*rc.get_edge((1, 0), d, rs, re).shape, dtype=dt, generator=g)
        a, b = oc.apply_edge((1, 0), d, st, env, Ve), rc.apply_edge((1, 0), d, rs, re, Ve)
        assert abs(complex(a) - complex(b)) < 1e-13 * abs(complex(b))
        ca = oc.corrf_1sO1sO((0, 0), d, st, env, op1, lambda r: ops2[r], 3)
        cb = rc.corrf_1sO1sO((0, 0), d, rs, re, op1, lambda r: ops2[r], 3)
        assert ca.shape == cb.shape and H.maxrel(ca, cb) < 1e-11, (d, ca, cb)
        # MPO pieces: rank-3 operator opens an MPO index, rank-4 passes it on, rank-3 closes it (corrf.py:423-450)
        o3, o4 = torch.randn(p, p, 3, dtype=dt, generator=g), torch.randn(p, p, 3, 2, dtype=dt, generator=g)
        o3b = torch.randn(p, p, 2, dtype=dt, generator=g)
        Ea, Eb = oc.apply_TM_1sO((0, 0), d, st, env, V, op=o3), rc.apply_TM_1sO((0, 0), d, rs, re, V, op=o3)
        assert Ea.shape == Eb.shape and H.maxrel(Ea, Eb) < 1e-13
        c1 = (d[0], d[1])
        Ea, Eb = oc.apply_TM_1sO(c1, d, st, env, Ea, op=o4), rc.apply_TM_1sO(c1, d, rs, re, Eb, op=o4)
        assert Ea.shape == Eb.shape and Ea.shape[-1] == 2 and H.maxrel(Ea, Eb) < 1e-12
        c2 = (2 * d[0], 2 * d[1])
        Ea, Eb = oc.apply_TM_1sO(c2, d, st, env, Ea, op=o3b), rc.apply_TM_1sO(c2, d, rs, re, Eb, op=o3b)
        assert Ea.dim() == 3 and Ea.shape == Eb.shape and H.maxrel(Ea, Eb) < 1e-12
        with pytest.raises(ValueError):
            oc.apply_TM_1sO((0, 0), d, st, env, V, op=o4)           # a rank-4 piece needs an incoming MPO index
        # two-site operators along the direction of growth
        t2 = torch.randn(p, p, p, p, dtype=dt, generator=g)
        t2s = [torch.randn(p, p, p, p, dtype=dt, generator=g) for _ in range(3)]
        for op in (None, t2):
            assert H.maxrel(oc.apply_TM_2sO_1sChannel((0, 0), d, st, env, V, op=op),
                            rc.apply_TM_2sO_1sChannel((0, 0), d, rs, re, V, op=op)) < 1e-12
        ca = oc.corrf_2sOH2sOH_E1((0, 0), d, st, env, t2, lambda r: t2s[r], 2)
        cb = rc.corrf_2sOH2sOH_E1((0, 0), d, rs, re, t2, lambda r: t2s[r], 2)
        assert H.maxrel(ca, cb) < 1e-10, (d, ca, cb)
        # width-2 edges, and (down / right only, as in the reference) the width-2 transfer matrix and its correlator
        assert H.maxrel(oc.get_edge_2((0, 0), d, st, env), rc.get_edge_2((0, 0), d, rs, re)) < 1e-13
        W = torch.randn(*rc.get_edge_2((0, 0), rev, rs, re).shape, dtype=dt, generator=g)
        We = torch.randn(*rc.get_edge_2((1, 1), d, rs, re).shape, dtype=dt, generator=g)
        sa, sb = oc.apply_edge((1, 1), d, st, env, We), rc.apply_edge((1, 1), d, rs, re, We)
        assert abs(complex(sa) - complex(sb)) < 1e-13 * abs(complex(sb))
        if d in ((0, 1), (1, 0)):
            for op in (None, t2):
                assert H.maxrel(oc.apply_TM_2sO_2sChannel((0, 0), d, st, env, W, op=op),
                                rc.apply_TM_2sO_2sChannel((0, 0), d, rs, re, W, op=op)) < 1e-12
            ca = oc.corrf_2sOV2sOV_E2((0, 0), d, st, env, t2, lambda r: t2s[r], 2)
            cb = rc.corrf_2sOV2sOV_E2((0, 0), d, rs, re, t2, lambda r: t2s[r], 2)
            assert H.maxrel(ca, cb) < 1e-10, (d, ca, cb)
        else:
            with pytest.raises(ValueError):
                oc.apply_TM_2sO_2sChannel((0, 0), d, st, env, W)
        # eigenvector edges in place of the environment's (the rl_0 argument)
        L = {c: torch.randn(*rc.get_edge(c, rev, rs, re).shape, dtype=dt, generator=g) for c in sites}
        R = {c: torch.randn(*rc.get_edge(c, d, rs, re).shape, dtype=dt, generator=g) for c in sites}
        rl_a = (lambda c: L[st.vertexToSite(c)], lambda c: R[st.vertexToSite(c)])
        rl_b = (lambda c: L[rs.vertexToSite(c)], lambda c: R[rs.vertexToSite(c)])
        ca = oc.corrf_1sO1sO((0, 0), d, st, env, op1, lambda r: ops2[r], 2, rl_0=rl_a)
        cb = rc.corrf_1sO1sO((0, 0), d, rs, re, op1, lambda r: ops2[r], 2, rl_0=rl_b)
        assert H.maxrel(ca, cb) < 1e-11, (d, ca, cb)


def c4v_case(name):
    """(site tensor, chi, C, T) of a C4v golden fixture (tests/golden/c4v_*.npz)."""
    z, meta = H.load_golden(name)
    return torch.from_numpy(z['site']), meta['chi'], torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_c4v_correlation_functions_match_reference(ref, name, monkeypatch):
    """ctm/one_site_c4v/corrf_c4v.py (get_edge, apply_edge, apply_TM_1sO, apply_TM_2sO, corrf_1sO1sO with and without rl_0,
    corrf_2sOH2sOH_E1) and transferops_c4v.get_Top_spec_c4v against the unmodified reference."""
    from ipeps.ipeps_c4v import IPEPS_C4V as RS
    from ctm.one_site_c4v.env_c4v import ENV_C4V as RE
    from ctm.one_site_c4v import corrf_c4v as rc, transferops_c4v as rt
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    from peps_torch_b200.ctm.one_site_c4v import corrf_c4v as oc, transferops_c4v as ot
    eng = H.OracleEngine()
    monkeypatch.setattr(oc, '_engine', lambda: eng)
    a, chi, C, T = c4v_case(name)
    dt = a.dtype
    ref.global_args.dtype, ref.global_args.torch_dtype, ref.global_args.device = ('complex128' if dt.is_complex else 'float64'), dt, 'cpu'
    rs, st = RS(a.clone()), IPEPS_C4V(a)
    re, env = RE(chi, rs), ENV_C4V(chi, st)
    for e in (re, env):
        e.C[e.keyC], e.T[e.keyT] = C.clone(), T.clone()
    p = a.shape[0]
    g = torch.Generator().manual_seed(9)
    op1 = torch.randn(p, p, dtype=dt, generator=g)
    ops1 = [torch.randn(p, p, dtype=dt, generator=g) for _ in range(4)]
    # two-site operators: real-valued (cast to the state's dtype), as the models' S.S is -- the reference's split of a complex
    # operator transposes V without conjugating it (corrf_c4v.py:470), which is reproduced but is not a decomposition of op
    op2 = torch.randn(p, p, p, p, dtype=torch.float64, generator=g).to(dt)
    ops2 = [torch.randn(p, p, p, p, dtype=torch.float64, generator=g).to(dt) for _ in range(3)]
    assert H.maxrel(oc.get_edge(st, env), rc.get_edge(rs, re)) < 1e-13
    V = torch.randn(chi, a.shape[1] ** 2, chi, dtype=dt, generator=g)
    assert abs(complex(oc.apply_edge(st, env, V)) - complex(rc.apply_edge(rs, re, V))) < 1e-13 * abs(complex(rc.apply_edge(rs, re, V)))
    for op in (None, op1):
        assert H.maxrel(oc.apply_TM_1sO(st, env, V, op=op), rc.apply_TM_1sO(rs, re, V, op=op)) < 1e-13
    for op in (None, op2):
        assert H.maxrel(oc.apply_TM_2sO(st, env, V, op=op), rc.apply_TM_2sO(rs, re, V, op=op)) < 1e-12
    ca, cb = oc.corrf_1sO1sO(st, env, op1, lambda r: ops1[r], 3), rc.corrf_1sO1sO(rs, re, op1, lambda r: ops1[r], 3)
    assert H.maxrel(ca, cb) < 1e-11, (ca, cb)
    L, R = torch.randn(*V.shape, dtype=dt, generator=g), torch.randn(*V.shape, dtype=dt, generator=g)
    ca = oc.corrf_1sO1sO(st, env, op1, lambda r: ops1[r], 2, rl_0=(L, R))
    cb = rc.corrf_1sO1sO(rs, re, op1, lambda r: ops1[r], 2, rl_0=(L, R))
    assert H.maxrel(ca, cb) < 1e-11, (ca, cb)
    ca, cb = oc.corrf_2sOH2sOH_E1(st, env, op2, lambda r: ops2[r], 2), rc.corrf_2sOH2sOH_E1(rs, re, op2, lambda r: ops2[r], 2)
    assert H.maxrel(ca, cb) < 1e-10, (ca, cb)
    La, Lb = ot.get_Top_spec_c4v(3, st, env), rt.get_Top_spec_c4v(3, rs, re)
    assert float((La - Lb).abs().max()) < 1e-10, (La, Lb)
    Ua, Ub = ot.get_Top_spec_c4v(2, st, env, normalize=False), rt.get_Top_spec_c4v(2, rs, re, normalize=False)
    assert float((Ua - Ub).abs().max()) < 1e-10 * float(Ub.abs().max()), (Ua, Ub)
    # width-2 edges and transfer matrix (vertical dimers, get_Top2_spec_c4v)
    for l in (1, 2, 3):
        assert H.maxrel(oc.get_edge_L(st, env, l=l), rc.get_edge_L(rs, re, l=l)) < 1e-13
    d2 = a.shape[1] ** 2
    W = torch.randn(chi, d2, d2, chi, dtype=dt, generator=g)
    sa, sb = oc.apply_edge_L(st, env, W), rc.apply_edge_L(rs, re, W)
    assert abs(complex(sa) - complex(sb)) < 1e-13 * abs(complex(sb))
    for op in (None, op2):
        assert H.maxrel(oc.apply_TM_1sO_2(st, env, W, op=op), rc.apply_TM_1sO_2(rs, re, W, op=op)) < 1e-12
    ca, cb = oc.corrf_2sOV2sOV_E2(st, env, op2, lambda r: ops2[r], 2), rc.corrf_2sOV2sOV_E2(rs, re, op2, lambda r: ops2[r], 2)
    assert H.maxrel(ca, cb) < 1e-10, (ca, cb)
    La, Lb = ot.get_Top2_spec_c4v(2, st, env), rt.get_Top2_spec_c4v(2, rs, re)
    assert float((La.abs().sort(0)[0] - Lb.abs().sort(0)[0]).abs().max()) < 1e-9, (La, Lb)


def _spec(S):
    """Magnitudes of an (n x 2) [re, im] list of eigenvalues, relative to the largest, ascending: the functions divide by the
    FIRST value in ARPACK's (reversed) order, which is not always the largest, and ARPACK's order depends on rounding."""
    m = (S[:, 0] ** 2 + S[:, 1] ** 2).sqrt()
    return (m / m.max()).sort()[0]


def test_entanglement_spectra_match_reference(ref, monkeypatch):
    """get_EH_spec_Ttensor of ctm/generic/transferops.py:207-370 (one-site cell: the kagome fixture, four directions) and of
    ctm/one_site_c4v/transferops_c4v.py:119-181: ring MPO of T tensors under ARPACK."""
    from ipeps.ipeps import IPEPS as RI
    from ctm.generic.env import ENV as RE
    from ctm.generic import transferops as rt
    from ipeps.ipeps_c4v import IPEPS_C4V as RS4
    from ctm.one_site_c4v.env_c4v import ENV_C4V as RE4
    from ctm.one_site_c4v import transferops_c4v as rt4
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    from peps_torch_b200.ctm.generic import transferops as ot, corrf as oc
    from peps_torch_b200.ctm.one_site_c4v import transferops_c4v as ot4
    eng = H.OracleEngine()
    monkeypatch.setattr(oc, '_engine', lambda: eng)
    z, meta = H.load_golden('kagome_1site_D2_chi8_A')
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'final_' if any(k.startswith('final_') for k in z.files) else 'mid_')
    dt = next(iter(sites.values())).dtype
    ref.global_args.dtype, ref.global_args.torch_dtype, ref.global_args.device = ('complex128' if dt.is_complex else 'float64'), dt, 'cpu'
    rs = RI(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    re = RE(meta['chi'], rs)
    re.C, re.T = dict(C), dict(T)
    st, env = H.State(sites, v2s, lX, lY), H.Env(meta['chi'], dict(C), dict(T))
    for d in DIRS:
        for L in (3, 4):
            Sa, Sb = ot.get_EH_spec_Ttensor(2, L, (0, 0), d, st, env), rt.get_EH_spec_Ttensor(2, L, (0, 0), d, rs, re)
            assert float((_spec(Sa) - _spec(Sb)).abs().max()) < 1e-9, (d, L, Sa, Sb)
    with pytest.warns(RuntimeWarning):
        assert ot.get_EH_spec_Ttensor(4, 2, (0, 0), (0, -1), st, env) is None         # D^L <= n
    for name in ('c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'):
        a, chi, C4, T4 = c4v_case(name)
        ref.global_args.dtype, ref.global_args.torch_dtype = ('complex128' if a.is_complex() else 'float64'), a.dtype
        rs4, st4 = RS4(a.clone()), IPEPS_C4V(a)
        re4, env4 = RE4(chi, rs4), ENV_C4V(chi, st4)
        for e in (re4, env4):
            e.C[e.keyC], e.T[e.keyT] = C4.clone(), T4.clone()
        for L in (3, 5):
            Sa, Sb = ot4.get_EH_spec_Ttensor(2, L, st4, env4), rt4.get_EH_spec_Ttensor(2, L, rs4, re4)
            assert float((_spec(Sa) - _spec(Sb)).abs().max()) < 1e-9, (name, L, Sa, Sb)
