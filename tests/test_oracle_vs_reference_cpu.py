"""Oracle restatements that have no committed fixture are pinned here against the reference itself.
Needs the reference tree (build container only): skipped on the GPU box."""
import os
import sys
from collections import OrderedDict
import pytest
import torch

REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ctm', 'generic')), reason='reference tree not present')


@pytest.fixture()
def ref(tmp_path):
    """The unmodified reference, importable, with cwd in a scratch directory (config.configure may write logs)
    and its CTMARGS singleton restored afterwards."""
    import copy
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
        cfg.global_args.dtype = 'float64'
        os.chdir(cwd)
        sys.path.remove(REF)


def _abs_worst(C1, T1, C2, T2):
    return max([float((C1[k].abs() - C2[k].abs()).abs().max()) for k in C2]
               + [float((T1[k].abs() - T2[k].abs()).abs().max()) for k in T2])


def _golden(name):
    import helpers as H
    z, meta = H.load_golden(name)
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'mid_')
    return sites, v2s, lX, lY, C0, T0, meta['chi']


def _ref_moves(cfg, sites, v2s, lX, lY, C0, T0, chi, oracle_args, ref_sites=None):
    """One reference ctm_MOVE per direction from (C0, T0) against the oracle's; |.| because fix_svd_signs can flip
    a column between two evaluations that differ by 1e-14."""
    import ctm_oracle as orc
    from ipeps.ipeps import IPEPS
    from ctm.generic.env import ENV
    from ctm.generic import ctmrg
    rs = ref_sites if ref_sites is not None else sites
    state = IPEPS(sites={c: t.clone() for c, t in rs.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    worst = 0.0
    for d in orc.DIRECTIONS:
        env = ENV(chi, state)
        env.C = {k: v.clone() for k, v in C0.items()}
        env.T = {k: v.clone() for k, v in T0.items()}
        ctmrg.ctm_MOVE(d, state, env, ctm_args=cfg.ctm_args, global_args=cfg.global_args)
        C, T = dict(C0), dict(T0)
        orc.ctm_move(d, sites, v2s, C, T, chi, oracle_args)
        worst = max(worst, _abs_worst(env.C, env.T, C, T))
    return worst


def test_oracle_4x2_projector_move_matches_reference(ref):
    import ctm_oracle as orc
    sites, v2s, lX, lY, C0, T0, chi = _golden('generic_4site_D3_chi12_B')
    ref.ctm_args.projector_method = '4X2'
    assert _ref_moves(ref, sites, v2s, lX, lY, C0, T0, chi, orc.OracleArgs(projector_method='4X2')) < 1e-11


def test_oracle_2norm_normalisation_matches_reference(ref):
    """ctm_absorb_normalization != 'inf' is the vector 2-norm (ctmrg.py:210-230)."""
    import ctm_oracle as orc
    sites, v2s, lX, lY, C0, T0, chi = _golden('generic_4site_D3_chi12_B')
    ref.ctm_args.ctm_absorb_normalization = 'fro'
    # (same rounding through S^-1/2 as in the other tests, 1e-11 at the threshold: SURVEY 8c quotes 7.7e-13 relative)
    assert _ref_moves(ref, sites, v2s, lX, lY, C0, T0, chi, orc.OracleArgs(ctm_absorb_normalization='fro')) < 1e-10


@pytest.mark.parametrize('name', ['generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128'])
def test_oracle_double_layer_move_matches_reference(ref, name):
    """Rank-4 (double-layer) on-site tensors, as ctmrg.run builds them under ctm_force_dl (ctmrg.py:51-61): the
    reference's dl branches of c2x2_* / absorb_truncate_* against the oracle's one-operand evaluation."""
    import ctm_oracle as orc
    sites, v2s, lX, lY, C0, T0, chi = _golden(name)
    dl = OrderedDict((c, orc.double_layer(a)) for c, a in sites.items())
    # the projector step dispatches on ctm_args.ctm_force_dl (ctm_projectors.py:50), the absorption on the rank of the
    # on-site tensor (ctmrg.py:342-346)
    ref.ctm_args.ctm_force_dl = True
    if sites[(0, 0)].is_complex():
        ref.global_args.dtype = 'complex128'
    # deterministic part element-wise: the halves of the reference's dl branch against the oracle's
    from ipeps.ipeps import IPEPS
    from ctm.generic.env import ENV
    from ctm.generic import ctm_components as cc
    st = IPEPS(sites=dict(dl), vertexToSite=v2s, lX=lX, lY=lY)
    env = ENV(chi, st)
    env.C, env.T = dict(C0), dict(T0)
    for d, f in zip(orc.DIRECTIONS, (cc.halves_of_4x4_CTM_MOVE_UP, cc.halves_of_4x4_CTM_MOVE_LEFT,
                                     cc.halves_of_4x4_CTM_MOVE_DOWN, cc.halves_of_4x4_CTM_MOVE_RIGHT)):
        for c in sites:
            R, Rt = f(c, st, env, mode='dl')
            R2, Rt2 = orc.halves(d, c, dl, v2s, C0, T0)
            assert float((R - R2).abs().max()) < 1e-13 * float(R.abs().max())
            assert float((Rt - Rt2).abs().max()) < 1e-13 * float(Rt.abs().max())
    # full moves: behind the SVD a 1e-16 difference of M is amplified by S0/Sj of the kept triplets (7e-6 in the complex
    # fixture => 1.4e-10 measured; SURVEY 8c: the reference reproduces itself to 4e-10 when the LAPACK driver changes)
    assert _ref_moves(ref, dl, v2s, lX, lY, C0, T0, chi, orc.OracleArgs()) < 1e-9
    # and the double-layer evaluation agrees with the single-layer one (same network, different contraction order)
    for d in orc.DIRECTIONS:
        C1, T1, C2, T2 = dict(C0), dict(T0), dict(C0), dict(T0)
        orc.ctm_move(d, sites, v2s, C1, T1, chi)
        orc.ctm_move(d, dl, v2s, C2, T2, chi)
        assert _abs_worst(C1, T1, C2, T2) < 1e-9


def test_oracle_overlap_tensor_matches_reference_run_overlap(ref):
    """run_overlap (ctmrg.py:112-175): ket = state1, bra = state2; one move per direction and iteration."""
    import ctm_oracle as orc
    from ipeps.ipeps import IPEPS
    from ctm.generic.env import ENV
    from ctm.generic import ctmrg
    sites, v2s, lX, lY, C0, T0, chi = _golden('generic_4site_D2_chi8_B')
    g = torch.Generator().manual_seed(7)
    sites2 = OrderedDict((c, a + 0.05 * (torch.rand(a.shape, generator=g, dtype=a.dtype) - 0.5)) for c, a in sites.items())
    s1 = IPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    s2 = IPEPS(sites={c: t.clone() for c, t in sites2.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    env = ENV(chi, s1)
    env.C = {k: v.clone() for k, v in C0.items()}
    env.T = {k: v.clone() for k, v in T0.items()}
    ref.ctm_args.ctm_force_dl = True
    ref.ctm_args.ctm_max_iter = 2
    ctmrg.run_overlap(s1, s2, env, ctm_args=ref.ctm_args, global_args=ref.global_args)
    dl = OrderedDict((c, orc.double_layer(sites[c], sites2[c])) for c in sites)
    C, T = dict(C0), dict(T0)
    for _ in range(2):
        for d in orc.DIRECTIONS:
            orc.ctm_move(d, dl, v2s, C, T, chi)
    assert _abs_worst(env.C, env.T, C, T) < 1e-10


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_oracle_c4v_double_layer_move_matches_reference(ref, name):
    """ctm_MOVE_dl (ctmrg_c4v.py:200-322) with truncated_eig_sym against the oracle's move on A = a (x) a*."""
    import helpers as H
    import ctm_oracle as orc
    from ctm.one_site_c4v import ctmrg_c4v
    from ctm.one_site_c4v.env_c4v import ENV_C4V
    from ipeps.ipeps_c4v import IPEPS_C4V
    from linalg.custom_eig import truncated_eig_sym
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    chi = meta['chi']
    if a.is_complex():
        ref.global_args.dtype = 'complex128'
    C, T = orc.init_env_c4v(a, chi)
    state = IPEPS_C4V(a.clone())
    env = ENV_C4V(chi, state)
    env.C[env.keyC], env.T[env.keyT] = C.clone(), T.clone()
    A = orc.double_layer(a)
    for _ in range(3):
        ctmrg_c4v.ctm_MOVE_dl(A, env, lambda M, chi_: truncated_eig_sym(M, chi_, keep_multiplets=True),
                              ctm_args=ref.ctm_args, global_args=ref.global_args)
        C, T = orc.ctm_move_c4v(A, C, T, chi)
        assert float((env.C[env.keyC] - C).abs().max()) < 1e-12
        assert float((env.T[env.keyT].abs() - T.abs()).abs().max()) < 1e-10


def test_oracle_rdm2x2_matches_reference_elementwise(ref):
    """rdm2x2_legacy (ctm/generic/rdm.py:1362-1592) element-wise, plus the open_sites semantics (partial traces)."""
    import helpers as H
    import ctm_oracle as orc
    from ipeps.ipeps import IPEPS
    from ctm.generic.env import ENV
    from ctm.generic import rdm
    for name in ('generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128'):
        z, meta = H.load_golden(name)
        sites = H.golden_sites(z)
        v2s, lX, lY = H.v2s_for(sites)
        C, T = H.golden_env(z, 'final_') if 'final_C_00_-1_-1' in z.files else H.golden_env(z, 'mid_')
        ref.global_args.dtype = 'complex128' if sites[(0, 0)].is_complex() else 'float64'
        st = IPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
        env = ENV(meta['chi'], st)
        env.C, env.T = dict(C), dict(T)
        for coord in sites:
            for spd in (False, True):
                r_ref = rdm.rdm2x2_legacy(coord, st, env, sym_pos_def=spd)
                r_orc = orc.rdm2x2(coord, sites, v2s, C, T, sym_pos_def=spd)
                assert float((r_ref - r_orc).abs().max()) < 1e-13
        raw = orc.rdm2x2((0, 0), sites, v2s, C, T, raw=True)
        tr = torch.einsum('ijklijkl', raw)
        nn = orc.rdm2x2((0, 0), sites, v2s, C, T, raw=True, open_sites=(0, 1))
        assert float((nn - torch.einsum('ijklabkl->ijab', raw)).abs().max()) < 1e-13 * abs(tr)
        nnn = orc.rdm2x2((0, 0), sites, v2s, C, T, raw=True, open_sites=(0, 3))
        assert float((nnn - torch.einsum('ijklajkd->ilad', raw)).abs().max()) < 1e-13 * abs(tr)
        one = orc.rdm2x2((0, 0), sites, v2s, C, T, raw=True, open_sites=(2,))
        assert float((one - torch.einsum('ijklijcl->kc', raw)).abs().max()) < 1e-13 * abs(tr)


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_oracle_c4v_rdms_match_reference_elementwise(ref, name):
    """rdm2x2_NN_lowmem_sl / rdm2x2_NNN_lowmem_sl / rdm2x2 of ctm/one_site_c4v/rdm_c4v.py against the generic
    construction on the rotated environment."""
    import helpers as H
    import ctm_oracle as orc
    from ctm.one_site_c4v import rdm_c4v
    from ctm.one_site_c4v.env_c4v import ENV_C4V
    from ipeps.ipeps_c4v import IPEPS_C4V
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    ref.global_args.dtype = 'complex128' if a.is_complex() else 'float64'
    C, T = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    st = IPEPS_C4V(a.clone())
    env = ENV_C4V(meta['chi'], st)
    env.C[env.keyC], env.T[env.keyT] = C.clone(), T.clone()
    for spd in (False, True):
        assert float((rdm_c4v.rdm2x2_NN_lowmem_sl(st, env, sym_pos_def=spd)
                      - orc.rdm2x2_c4v(a, C, T, (0, 1), spd)).abs().max()) < 1e-12
        assert float((rdm_c4v.rdm2x2_NNN_lowmem_sl(st, env, sym_pos_def=spd)
                      - orc.rdm2x2_c4v(a, C, T, (0, 3), spd)).abs().max()) < 1e-12
        assert float((rdm_c4v.rdm2x2(st, env, sym_pos_def=spd) - orc.rdm2x2_c4v(a, C, T, (0, 1, 2, 3), spd)).abs().max()) < 1e-12


def test_oracle_small_rdms_match_reference_elementwise(ref):
    """rdm1x1_dl / rdm2x1_dl / rdm1x2_dl (ctm/generic/rdm.py:114-258,352-500,672-826) element-wise."""
    import helpers as H
    import ctm_oracle as orc
    from ipeps.ipeps import IPEPS
    from ctm.generic.env import ENV
    from ctm.generic import rdm
    for name in ('generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128', 'generic_4site_D3_chi12_B'):
        z, meta = H.load_golden(name)
        sites = H.golden_sites(z)
        v2s, lX, lY = H.v2s_for(sites)
        C, T = H.golden_env(z, 'mid_')
        ref.global_args.dtype = 'complex128' if sites[(0, 0)].is_complex() else 'float64'
        st = IPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
        env = ENV(meta['chi'], st)
        env.C, env.T = dict(C), dict(T)
        for coord in sites:
            for spd in (False, True):
                for f_ref, f_orc in ((rdm.rdm1x1_dl, orc.rdm1x1), (rdm.rdm2x1_dl, orc.rdm2x1), (rdm.rdm1x2_dl, orc.rdm1x2)):
                    r_ref = f_ref(coord, st, env, sym_pos_def=spd)
                    r_orc = f_orc(coord, sites, v2s, C, T, sym_pos_def=spd)
                    assert r_ref.shape == r_orc.shape
                    assert float((r_ref - r_orc).abs().max()) < 1e-13, (name, coord, spd, f_orc.__name__)
            # operator= : the unnormalised expectation value (rdm.py:89-90,175-181) through the host wrapper over the raw network
            from peps_torch_b200.ctm.generic import rdm as ours
            p = sites[coord].shape[0]
            op = torch.randn(p, p, dtype=sites[coord].dtype, generator=torch.Generator().manual_seed(3))
            saved, ours._engine = ours._engine, (lambda: H.OracleEngine())
            try:
                got = ours.rdm1x1(coord, H.State(sites, v2s, lX, lY), H.Env(meta['chi'], dict(C), dict(T)), operator=op)
            finally:
                ours._engine = saved
            want = rdm.rdm1x1_dl(coord, st, env, operator=op)
            assert abs(complex(got) - complex(want)) < 1e-13 * abs(complex(want)), (name, coord, got, want)


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_oracle_c4v_small_rdms_match_reference_elementwise(ref, name):
    import helpers as H
    import ctm_oracle as orc
    from ctm.one_site_c4v import rdm_c4v
    from ctm.one_site_c4v.env_c4v import ENV_C4V
    from ipeps.ipeps_c4v import IPEPS_C4V
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    ref.global_args.dtype = 'complex128' if a.is_complex() else 'float64'
    C, T = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    st = IPEPS_C4V(a.clone())
    env = ENV_C4V(meta['chi'], st)
    env.C[env.keyC], env.T[env.keyT] = C.clone(), T.clone()
    for spd in (False, True):
        for f_ref, kind in ((rdm_c4v.rdm1x1_sl, '1x1'), (rdm_c4v.rdm1x1, '1x1'), (rdm_c4v.rdm2x1_sl, '2x1'), (rdm_c4v.rdm2x1, '2x1')):
            r_ref = f_ref(st, env, sym_pos_def=spd)
            r_orc = orc.rdm_small_c4v(kind, a, C, T, spd)
            assert r_ref.shape == r_orc.shape
            assert float((r_ref - r_orc).abs().max()) < 1e-12, (kind, spd, f_ref.__name__)
        # rdm3x1(_sl) (rdm_c4v.py:667-1011): the J3 term of energy_1x1_lowmem (models/j1j2.py:671-677)
        for f_ref in (rdm_c4v.rdm3x1_sl, rdm_c4v.rdm3x1):
            r_ref, r_orc = f_ref(st, env, sym_pos_def=spd), orc.rdm3x1_c4v(a, C, T, spd)
            assert r_ref.shape == r_orc.shape and float((r_ref - r_orc).abs().max()) < 1e-13, (spd, f_ref.__name__)


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_oracle_c4v_qr_move_matches_reference(ref, name, monkeypatch):
    """ctm_MOVE_QR_sl (ctmrg_c4v.py:465-602): the oracle's restatement element-wise (the same LAPACK QR on both sides), and
    the drop-in composed from engine calls (peps_torch_b200/ctm/one_site_c4v/ctmrg_c4v.py, oracle as engine) through |.|."""
    import helpers as H
    import ctm_oracle as orc
    from ctm.one_site_c4v import ctmrg_c4v
    from ctm.one_site_c4v.env_c4v import ENV_C4V
    from ipeps.ipeps_c4v import IPEPS_C4V
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v as ours
    from peps_torch_b200.env import ENV_C4V as OurEnv
    from peps_torch_b200.ipeps import IPEPS_C4V as OurState
    from peps_torch_b200.config import CTMARGS
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    chi = meta['chi']
    if a.is_complex():
        ref.global_args.dtype = 'complex128'
    C, T = orc.init_env_c4v(a, chi)
    for _ in range(2):
        C, T = orc.ctm_move_c4v(a, C, T, chi)
    state = IPEPS_C4V(a.clone())
    env = ENV_C4V(chi, state)
    env.C[env.keyC], env.T[env.keyT] = C.clone(), T.clone()
    eng = H.OracleEngine()
    monkeypatch.setattr(ours, '_engine', lambda: eng)
    env2 = OurEnv(chi, OurState(a.clone()))
    env2.C[env2.keyC], env2.T[env2.keyT] = C.clone(), T.clone()
    for _ in range(3):
        ctmrg_c4v.ctm_MOVE_QR_sl(a, env, ctm_args=ref.ctm_args, global_args=ref.global_args)
        C, T = orc.ctm_move_qr_c4v(a, C, T, chi)
        ours.ctm_MOVE_QR_sl(a, env2, ctm_args=CTMARGS())
        assert float((env.C[env.keyC] - C).abs().max()) < 1e-12
        assert float((env.T[env.keyT] - T).abs().max()) < 1e-12
        assert float((env2.C[env2.keyC].abs() - C.abs()).abs().max()) < 1e-12
        assert float((env2.T[env2.keyT].abs() - T.abs()).abs().max()) < 1e-12


@pytest.mark.parametrize('dt', [torch.float64, torch.complex128])
def test_oracle_rectangular_cell_direction_dependent_bonds_matches_reference_run(ref, dt):
    """3 x 2 unit cell, six different sites, vertical bonds D = 2 and horizontal bonds D = 3: the reference's own run
    (ctm/generic/ctmrg.py:18-110, init_env 'CTMRG') for two iterations against the oracle's -- pins the oracle for the
    GPU test test_rectangular_cell_with_direction_dependent_bonds."""
    import helpers as H
    import ctm_oracle as orc
    from ipeps.ipeps import IPEPS
    from ctm.generic.env import ENV, init_env
    from ctm.generic import ctmrg
    lX, lY, chi, Dv, Dh = 3, 2, 12, 2, 3
    g = torch.Generator().manual_seed(77)
    sites = OrderedDict()
    for y in range(lY):
        for x in range(lX):
            a = torch.randn(2, Dv, Dh, Dv, Dh, dtype=dt, generator=g)
            sites[(x, y)] = a / a.abs().max()

    def v2s(c):
        return (c[0] % lX, c[1] % lY)
    C, T = orc.init_env(sites, v2s, chi)
    orc.run(sites, v2s, lX, lY, C, T, chi, 2)
    ref.global_args.dtype, ref.global_args.torch_dtype = ('complex128' if dt.is_complex else 'float64'), dt
    ref.ctm_args.ctm_max_iter, ref.ctm_args.projector_svd_method = 2, 'GESDD'
    try:
        state = IPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
        env = ENV(chi, state)
        init_env(state, env)
        env, *_ = ctmrg.run(state, env, conv_check=None, ctm_args=ref.ctm_args)
    finally:
        ref.global_args.torch_dtype = torch.float64
    assert H.spectra_diff(env.C, C) < 1e-10
    assert _abs_worst(env.C, env.T, C, T) < 4e-9       # the reference-vs-reference floor of SURVEY 8c
