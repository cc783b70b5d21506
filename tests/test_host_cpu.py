"""CPU suite, host side: the C-ABI library loads and exports every symbol include/ctmb.h
declares; error reporting works without a GPU; site tables agree with the oracle's."""
import os
import re
import ctypes
import pytest
import torch
import ctm_oracle as orc
import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from peps_torch_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'ctmb.h')).read()
    declared = set(re.findall(r'\b(ctmb_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(_lib.lib, name), f'{name} declared in ctmb.h but not exported'
    assert declared == set(_lib.SIGNATURES), 'ctypes signatures out of sync with ctmb.h'
    assert _lib.lib.ctmb_version() >= 100


def test_struct_layout_matches_header():
    from peps_torch_b200 import _lib
    assert ctypes.sizeof(_lib.Options) == 72      # 4 doubles, 4 ints, u64, double, 2 ints (include/ctmb.h)
    assert ctypes.sizeof(_lib.Site) == 8 + 24 + 32 + 32
    o = _lib.default_options()
    assert (o.svd_reltol, o.eps_multiplet, o.multiplet_abstol) == (1e-8, 1e-8, 1e-14)
    assert o.rsvd_niter >= 2 and o.rsvd_rank_factor == 0.0      # 0 = automatic sketch width
    assert o.projector_method == 0 and 0.0 < o.rsvd_tol < 1e-13


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU error path')
def test_no_cpu_fallback():
    from peps_torch_b200 import _lib
    from peps_torch_b200.engine import CtmEngine
    h = ctypes.c_void_p()
    assert _lib.lib.ctmb_create(ctypes.byref(h), 0) != 0
    assert b'CUDA' in _lib.lib.ctmb_last_error()
    with pytest.raises(RuntimeError):
        CtmEngine()
    from peps_torch_b200.ctm.generic import ctmrg
    with pytest.raises(RuntimeError):
        ctmrg.ctm_MOVE((0, -1), None, None)
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    with pytest.raises(RuntimeError):
        ctmrg_c4v.ctm_MOVE_sl(None, None)


def test_move_tables_agree_with_oracle():
    from peps_torch_b200 import engine as E
    for d in orc.DIRECTIONS:
        flat = [(k, dx, dy, tr) for pair in orc.HALVES[d] for (k, dx, dy, tr) in pair]
        assert [(dx, dy) for (_, dx, dy, _) in flat] == E.PATCH[d]
        assert orc.ABSORB[d]['shift'] == E.SHIFT[d]
        assert orc.ABSORB[d]['out'] == E.OUT_KEYS[d]
    assert [orc.CORNERS[k][0] for k in ('LU', 'RU', 'RD', 'LD')] == E.C_KEYS


def test_host_env_init_matches_oracle_and_reference_fixture():
    from peps_torch_b200.ipeps import IPEPS, IPEPS_C4V
    from peps_torch_b200.env import ENV, init_env, ENV_C4V, init_env_c4v
    z, meta = H.load_golden('generic_4site_D2_chi8_B')
    sites = H.golden_sites(z)
    st = IPEPS(sites, orc.v2s_4site, 2, 2)
    env = ENV(meta['chi'], st)
    init_env(st, env)
    Cg, Tg = H.golden_env(z, 'init_')
    assert all(torch.equal(env.C[k], Cg[k]) for k in Cg) and all(torch.equal(env.T[k], Tg[k]) for k in Tg)
    z, meta = H.load_golden('c4v_D2_chi8_B')
    st = IPEPS_C4V(torch.from_numpy(z['site']))
    env = ENV_C4V(meta['chi'], st)
    init_env_c4v(st, env)
    assert H.maxrel(env.get_C(), torch.from_numpy(z['init_C'])) < 1e-14
    assert H.maxrel(env.get_T().abs(), torch.from_numpy(z['init_T']).abs()) < 1e-12


@pytest.fixture()
def oracle_engine(monkeypatch):
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    e = H.OracleEngine()
    monkeypatch.setattr(ctmrg, '_engine', lambda: e)
    monkeypatch.setattr(ctmrg_c4v, '_engine', lambda: e)
    return e


def test_host_run_variants_control_flow(oracle_engine):
    """ctm_force_dl, warm-up, run_overlap (ctmrg.py:51-61,76-86,112-175): host logic with the oracle as engine."""
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS
    z, meta = H.load_golden('generic_4site_D2_chi8_B')
    chi = meta['chi']
    sites = H.golden_sites(z)
    C0, T0 = H.golden_env(z, 'init_')
    st = IPEPS(sites, orc.v2s_4site, 2, 2)
    C, T = dict(C0), dict(T0)
    for _ in range(3):
        orc.ctm_iteration(sites, orc.v2s_4site, 2, 2, C, T, chi)
    # ctm_force_dl: every move sees rank-4 tensors, result = single-layer result
    args = CTMARGS(); args.ctm_force_dl = True; args.ctm_max_iter = 3
    env = H.Env(chi, dict(C0), dict(T0))
    ctmrg.run(st, env, ctm_args=args)
    assert len(oracle_engine.calls) == 24 and all(c[2] == 4 for c in oracle_engine.calls)
    assert H.env_abs_diff(env.C, env.T, C, T) < 1e-9
    # warm-up: max(ctm_warmup_iter, ceil(chi / D^2)) = 2 iterations before the ctm_max_iter = 1 of the main loop
    oracle_engine.calls.clear()
    args = CTMARGS(); args.ctm_warmup_iter = 1; args.ctm_max_iter = 1
    env = H.Env(chi, dict(C0), dict(T0))
    seen = []
    ctmrg.run(st, env, conv_check=lambda s, e, h, ctm_args=None: (seen.append(1) or False, h), ctm_args=args)
    assert len(oracle_engine.calls) == 24 and all(c[2] == 5 for c in oracle_engine.calls) and len(seen) == 1
    assert H.env_abs_diff(env.C, env.T, C, T) < 1e-12
    # option mapping
    a2 = CTMARGS(); a2.ctm_absorb_normalization = 'fro'; a2.projector_method = '4X2'; a2.projector_svd_method = 'ARP'
    o = ctmrg._options(a2)
    assert (o['norm_type'], o['projector_method']) == (1, 1)
    # run_overlap: needs ctm_force_dl, one move per direction and iteration, conv_check(state1, state2, env, history)
    oracle_engine.calls.clear()
    args = CTMARGS(); args.ctm_max_iter = 2
    with pytest.raises(AssertionError):
        ctmrg.run_overlap(st, st, env, ctm_args=args)
    args.ctm_force_dl = True
    env = H.Env(chi, dict(C0), dict(T0))
    ctmrg.run_overlap(st, st, env, conv_check=lambda s1, s2, e, h, ctm_args=None: (False, h), ctm_args=args)
    assert len(oracle_engine.calls) == 8 and all(c[2] == 4 for c in oracle_engine.calls)
    Co, To = dict(C0), dict(T0)
    for _ in range(2):
        for d in orc.DIRECTIONS:
            orc.ctm_move(d, sites, orc.v2s_4site, Co, To, chi)
    assert H.env_abs_diff(env.C, env.T, Co, To) < 1e-9


def test_host_c4v_run_dl_control_flow(oracle_engine):
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    z, meta = H.load_golden('c4v_D2_chi8_B')
    chi = meta['chi']
    a = torch.from_numpy(z['site'])
    st = IPEPS_C4V(a)
    C, T = torch.from_numpy(z['init_C']), torch.from_numpy(z['init_T'])
    for _ in range(3):
        C, T = orc.ctm_move_c4v(a, C, T, chi)
    args = CTMARGS(); args.ctm_max_iter = 3
    for runner, rank in ((ctmrg_c4v.run, 5), (ctmrg_c4v.run_dl, 4)):
        oracle_engine.calls.clear()
        env = ENV_C4V(chi, st)
        env.C[env.keyC], env.T[env.keyT] = torch.from_numpy(z['init_C']), torch.from_numpy(z['init_T'])
        runner(st, env, ctm_args=args)
        assert oracle_engine.calls == [('c4v', rank)] * 3
        assert H.maxrel(env.get_C(), C) < 1e-10 and H.maxrel(env.get_T().abs(), T.abs()) < 1e-9
    bad = CTMARGS(); bad.projector_svd_method = 'GESDD'
    with pytest.raises(Exception):
        ctmrg_c4v.run(st, env, ctm_args=bad)


def test_planning_only_handle_validates_every_entry_point_and_never_computes():
    """ctmb_create(device=-1): the *_workspace queries run the planner of every entry point (all contraction labels and
    extents are validated) without a GPU; compute calls fail loudly."""
    import ctypes as C
    from peps_torch_b200 import _lib
    from peps_torch_b200.engine import CtmEngine, C_KEYS, T_KEYS, DIRECTIONS
    from peps_torch_b200.ipeps import IPEPS
    lib = _lib.lib
    eng = CtmEngine('plan')
    for name in ('generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'):
        z, meta = H.load_golden(name)
        chi = meta['chi']
        sites = H.golden_sites(z)
        v2s, lX, lY = H.v2s_for(sites)
        C0, T0 = H.golden_env(z, 'mid_')
        for dl in (False, True):
            ss = type(sites)((c, orc.double_layer(a)) for c, a in sites.items()) if dl else sites
            st = IPEPS(ss, v2s, lX, lY)
            env = H.Env(chi, C0, T0)
            keep = []
            coords, arr = eng._sites_array(st, env, keep)
            dt = _lib.F64 if not ss[coords[0]].is_complex() else _lib.C128
            for d, dcode in DIRECTIONS.items():
                corner, nb, dest, _ = eng._move_tables(st, d)
                for pm in (0, 1):
                    o = eng._opts(projector_method=pm)
                    assert lib.ctmb_move_generic_workspace(eng._h, dt, dcode, len(coords), chi, arr, corner, nb, C.byref(o)) > 0, \
                        lib.ctmb_last_error()
            for k in range(4):
                assert lib.ctmb_c2x2_workspace(eng._h, dt, k, chi, C.byref(arr[0])) > 0, lib.ctmb_last_error()
            # reduced density matrices: every subset of open sites (single-layer tensors only)
            ptrs = (C.POINTER(_lib.Site) * 4)(*[C.pointer(arr[i % len(coords)]) for i in range(4)])
            for mask in range(1, 16):
                nbytes = lib.ctmb_rdm2x2_workspace(eng._h, dt, chi, ptrs, mask)
                assert (nbytes > 0) == (not dl), (name, dl, mask, lib.ctmb_last_error())
            assert lib.ctmb_rdm2x2_workspace(eng._h, dt, chi, ptrs, 0) == 0
            two = (C.POINTER(_lib.Site) * 2)(C.pointer(arr[0]), C.pointer(arr[1 % len(coords)]))
            for kind in range(3):
                # the sites of the fixtures have equal bond dimensions, so any pair of them plans
                nbytes = lib.ctmb_rdm_small_workspace(eng._h, dt, kind, chi, two)
                assert (nbytes > 0) == (not dl), (name, dl, kind, lib.ctmb_last_error())
            # a compute call on the planning handle is an error, not a fallback
            with pytest.raises(_lib.CtmbError):
                eng.move_generic((0, -1), st, env)
    for n, spd in ((4, 1), (16, 1), (256, 0)):
        assert lib.ctmb_sym_pos_def_workspace(eng._h, _lib.C128, n, spd) > 0, lib.ctmb_last_error()
    assert lib.ctmb_sym_pos_def_workspace(eng._h, _lib.F64, 256, 1) == 0          # complete eigendecomposition: n <= 160
    for name in ('c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'):
        z, meta = H.load_golden(name)
        a = torch.from_numpy(z['site'])
        dt = _lib.C128 if a.is_complex() else _lib.F64
        for t in (a, orc.double_layer(a)):
            dims = (C.c_int * 5)(*(list(t.shape) if t.dim() == 5 else [0] + list(t.shape)))
            assert lib.ctmb_move_c4v_workspace(eng._h, dt, dims, meta['chi'], None) > 0, lib.ctmb_last_error()


def test_host_rdm_modules_map_sites_and_env_correctly(monkeypatch):
    """ctm/generic/rdm.py and ctm/one_site_c4v/rdm_c4v.py drop-ins: which site / which rotated env tensor goes where
    (host logic, oracle as engine)."""
    from peps_torch_b200.ctm.generic import rdm
    from peps_torch_b200.ctm.one_site_c4v import rdm_c4v
    from peps_torch_b200.ipeps import IPEPS, IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    e = H.OracleEngine()
    monkeypatch.setattr(rdm, '_engine', lambda: e)
    monkeypatch.setattr(rdm_c4v, '_engine', lambda: e)
    z, meta = H.load_golden('generic_4site_D2_chi8_B')
    sites = H.golden_sites(z)
    C, T = H.golden_env(z, 'mid_')
    st = IPEPS(sites, orc.v2s_4site, 2, 2)
    env = H.Env(meta['chi'], C, T)
    r = rdm.rdm2x2((1, 0), st, env)
    assert r.shape == (2,) * 8 and abs(float(torch.einsum('ijklijkl', r)) - 1.0) < 1e-13
    assert torch.equal(r, orc.rdm2x2((1, 0), sites, orc.v2s_4site, C, T))
    z, meta = H.load_golden('c4v_D2_chi8_B_c128')
    a = torch.from_numpy(z['site'])
    stc = IPEPS_C4V(a)
    envc = ENV_C4V(meta['chi'], stc)
    Cc, Tc = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    envc.C[envc.keyC], envc.T[envc.keyT] = Cc, Tc
    assert H.maxrel(rdm_c4v.rdm2x2_NN_lowmem_sl(stc, envc, sym_pos_def=True), orc.rdm2x2_c4v(a, Cc, Tc, (0, 1), True)) < 1e-13
    assert H.maxrel(rdm_c4v.rdm2x2_NNN_lowmem_sl(stc, envc), orc.rdm2x2_c4v(a, Cc, Tc, (0, 3))) < 1e-13
    assert H.maxrel(rdm_c4v.rdm2x2(stc, envc), orc.rdm2x2_c4v(a, Cc, Tc)) < 1e-13
    assert H.maxrel(rdm_c4v.rdm1x1_sl(stc, envc), orc.rdm_small_c4v('1x1', a, Cc, Tc)) < 1e-13
    assert H.maxrel(rdm_c4v.rdm2x1_sl(stc, envc, sym_pos_def=True), orc.rdm_small_c4v('2x1', a, Cc, Tc, True)) < 1e-13
    assert H.maxrel(rdm_c4v.rdm3x1_sl(stc, envc, sym_pos_def=True), orc.rdm3x1_c4v(a, Cc, Tc, True)) < 1e-13
    assert H.maxrel(rdm_c4v.rdm3x1(stc, envc), orc.rdm3x1_c4v(a, Cc, Tc)) < 1e-13
    for f, g in ((rdm.rdm1x1, orc.rdm1x1), (rdm.rdm2x1, orc.rdm2x1), (rdm.rdm1x2, orc.rdm1x2)):
        assert torch.equal(f((0, 1), st, env), g((0, 1), sites, orc.v2s_4site, C, T))
    # operator=: the unnormalised expectation value; with the identity, the trace of the raw network (rdm.py:89-90)
    raw = orc.rdm1x1((0, 0), sites, orc.v2s_4site, C, T, raw=True)
    val = rdm.rdm1x1((0, 0), st, env, operator=torch.eye(2, dtype=raw.dtype))
    assert val.dim() == 0 and abs(float(val) - float(raw.diagonal().sum())) < 1e-13 * abs(float(val))


def test_anisotropic_and_ragged_unit_cells_plan_or_fail_loudly():
    """Edge cases of the unit cell, checked with the planning-only handle: (1) bond dimensions that differ between the
    horizontal and the vertical bonds but are the same for every site plan in all four directions; (2) a consistent cell
    whose vertical bonds alternate D = 2 / 3 between the rows (the reference handles it, and so does the oracle): the
    moves along the uniform direction plan, the others are refused with an explicit message -- libctmb batches the site
    jobs of a move and needs equal projector shapes -- never a silent wrong result."""
    import ctypes as C
    from collections import OrderedDict
    from peps_torch_b200 import _lib
    from peps_torch_b200.engine import CtmEngine, DIRECTIONS
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    eng = CtmEngine('plan')
    g = torch.Generator().manual_seed(3)
    chi = 6

    def plan(shapes):
        sites = OrderedDict((c, torch.rand(s, dtype=torch.float64, generator=g) - 0.5) for c, s in shapes.items())
        st = IPEPS(sites, orc.v2s_4site, 2, 2)
        env = ENV(chi, st)
        init_env(st, env)
        Cc, Tc = orc.init_env(sites, orc.v2s_4site, chi)
        orc.ctm_iteration(sites, orc.v2s_4site, 2, 2, Cc, Tc, chi)          # the oracle (= the reference algorithm) runs
        keep = []
        coords, arr = eng._sites_array(st, env, keep)
        out = {}
        for d, dc in DIRECTIONS.items():
            corner, nb, dest, _ = eng._move_tables(st, d)
            o = eng._opts()
            n = _lib.lib.ctmb_move_generic_workspace(eng._h, _lib.F64, dc, 4, chi, arr, corner, nb, C.byref(o))
            out[d] = (n, _lib.lib.ctmb_last_error().decode())
        return out

    cells = [(0, 0), (1, 0), (0, 1), (1, 1)]
    res = plan({c: (2, 2, 3, 2, 3) for c in cells})
    assert all(n > 0 for n, _ in res.values()), res
    res = plan({(0, 0): (2, 2, 2, 3, 2), (1, 0): (2, 2, 2, 3, 2), (0, 1): (2, 3, 2, 2, 2), (1, 1): (2, 3, 2, 2, 2)})
    assert res[(0, -1)][0] > 0 and res[(0, 1)][0] > 0
    for d in ((-1, 0), (1, 0)):
        assert res[d][0] == 0 and 'non-uniform bond dimensions' in res[d][1]
