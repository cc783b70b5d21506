"""CPU suite: the oracle against the golden fixtures written from the reference
(oracle/gen_golden.py), the reference's own known answers, and the host-side logic.
No CUDA device is needed for anything in this file."""
import numpy as np
import pytest
import torch
import ctm_oracle as orc
import helpers as H

GENERIC = ['generic_4site_D2_chi8_A', 'generic_4site_D2_chi8_B', 'generic_4site_D3_chi12_B',
           'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A']
C4V = ['c4v_D2_chi8_A', 'c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128']


@pytest.mark.parametrize('name', GENERIC)
def test_oracle_pieces_match_reference_fixtures(name):
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    # initial environment (ctm/generic/env.py:367-536): bit-exact
    C0, T0 = orc.init_env(sites, v2s, chi)
    Cg, Tg = H.golden_env(z, 'init_')
    for k in Cg:
        assert torch.equal(C0[k], Cg[k])
    for k in Tg:
        assert torch.equal(T0[k], Tg[k])
    # deterministic pieces on the reference's environment after one iteration: 1e-13
    C, T = H.golden_env(z, 'mid_')
    coord = list(sites.keys())[-1]
    for kind in orc.CORNERS:
        o = orc.corner_at(kind, coord, sites, v2s, C, T)
        assert H.maxrel(o, torch.from_numpy(z['c2x2_' + kind])) < 1e-13
    args = orc.OracleArgs()
    for d in orc.DIRECTIONS:
        tg = f'{d[0]}_{d[1]}'
        R, Rt = orc.halves(d, coord, sites, v2s, C, T)
        assert H.maxrel(R, torch.from_numpy(z[f'halves_{tg}_R'])) < 1e-13
        assert H.maxrel(Rt, torch.from_numpy(z[f'halves_{tg}_Rt'])) < 1e-13
        # projectors from the reference's own R, Rt: gauge-invariant product
        Rr, Rtr = torch.from_numpy(z[f'halves_{tg}_R']), torch.from_numpy(z[f'halves_{tg}_Rt'])
        P, Pt = orc.projectors_from_matrices(Rr, Rtr, chi, args)
        Pr, Ptr = torch.from_numpy(z[f'proj_{tg}_P']), torch.from_numpy(z[f'proj_{tg}_Pt'])
        assert H.maxrel(P @ Pt.t(), Pr @ Ptr.t()) < 1e-9
        # whole move from the same snapshot: |C|,|T| (SVD sign gauge is not reproducible)
        C2 = {k: v.clone() for k, v in C.items()}
        T2 = {k: v.clone() for k, v in T.items()}
        orc.ctm_move(d, sites, v2s, C2, T2, chi, args)
        Cm, Tm = H.golden_env(z, f'move_{tg}_')
        assert H.env_abs_diff(C2, T2, Cm, Tm) < 1e-8


@pytest.mark.parametrize('name', [n for n in GENERIC if 'kagome' not in n])
def test_oracle_run_energy_and_spectra(name):
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'mid_')
    orc.run(sites, v2s, lX, lY, C, T, chi, meta['n_iter'])
    Cf, Tf = H.golden_env(z, 'final_')
    assert H.spectra_diff(C, Cf) < 1e-9
    assert H.env_abs_diff(C, T, Cf, Tf) < 1e-7
    e = orc.energy_j1j2(sites, v2s, C, T, 1.0, meta['j2'])
    assert abs(e - float(z['energy'][0])) < 1e-10 * max(1.0, abs(e))
    # rdm2x2 of the reference's final environment: deterministic, 1e-11
    rho = orc.rdm2x2(list(sites.keys())[-1], sites, v2s, Cf, Tf)
    assert H.maxrel(rho, torch.from_numpy(z['rdm2x2'])) < 1e-11


@pytest.mark.parametrize('name', C4V)
def test_oracle_c4v_matches_reference_fixtures(name):
    z, meta = H.load_golden(name)
    chi = meta['chi']
    a = torch.from_numpy(z['site'])
    C, T = orc.init_env_c4v(a, chi)
    assert H.maxrel(C, torch.from_numpy(z['init_C'])) < 1e-14
    assert H.maxrel(T.abs(), torch.from_numpy(z['init_T']).abs()) < 1e-12
    Cm, Tm = torch.from_numpy(z['mid_C']), torch.from_numpy(z['mid_T'])
    assert H.maxrel(orc.c2x2_c4v(a, Cm, Tm), torch.from_numpy(z['mid_c2x2'])) < 1e-13
    nC, nT = orc.ctm_move_c4v(a, Cm, Tm, chi)
    assert H.maxrel(nC, torch.from_numpy(z['mid_nC'])) < 1e-12
    assert H.maxrel(nT.abs(), torch.from_numpy(z['mid_nT']).abs()) < 1e-10
    Cf, Tf = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    e = orc.energy_j1j2_c4v(a, Cf, Tf, 1.0, meta['j2'])
    assert abs(e - float(z['energy'][0])) < 1e-11


def test_known_answer_rvb_c4v():
    """TestRVB.test_ctmrg_RVB of the reference (examples/j1j2/ctmrg_j1j2_c4v.py:214-259):
    e = -0.47684229 @1e-8 for RVB_1x1.in, chi=16, j2=0.5."""
    z = np.load(H.GOLD + '/rvb_c4v_known_answer.npz')
    a = torch.from_numpy(z['site'])
    chi = int(z['chi'][0])
    C, T = orc.init_env_c4v(a, chi)
    for _ in range(200):
        C, T = orc.ctm_move_c4v(a, C, T, chi)
    e = orc.energy_j1j2_c4v(a, C, T, 1.0, float(z['j2'][0]))
    assert abs(e - float(z['energy'][0])) < 1e-8


def test_known_answer_j1j2_2site():
    """TestCtmrg_States, 2SITE ansatz (examples/j1j2/ctmrg_j1j2.py:259-267): -0.4434603770143078 @1e-6."""
    z = np.load(H.GOLD + '/j1j2_2site_known_answer.npz')
    sites = H.golden_sites(z)
    chi, j2 = int(z['chi'][0]), float(z['j2'][0])
    C, T = orc.init_env(sites, orc.v2s_2site, chi)
    e_prev = None
    for _ in range(30):
        orc.ctm_iteration(sites, orc.v2s_2site, int(z['lX'][0]), int(z['lY'][0]), C, T, chi)
        e = orc.energy_j1j2(sites, orc.v2s_2site, C, T, 1.0, j2)
        if e_prev is not None and abs(e - e_prev) < 1e-8:
            break
        e_prev = e
    assert abs(e - float(z['energy'][0])) < 1e-6


def test_sl_einsum_equals_double_layer():
    torch.manual_seed(3)
    a = torch.rand(2, 2, 3, 2, 3, dtype=torch.complex128) - (0.5 + 0.5j)     # ragged bond dims
    chi = 5
    C = torch.rand(chi, chi, dtype=a.dtype)
    T1 = torch.rand(chi, 4, chi, dtype=a.dtype)
    T2 = torch.rand(chi, chi, 9, dtype=a.dtype)
    A = orc.double_layer(a)
    ref = torch.einsum('ab,buc,ael,ulfg->efcg', C, T1, T2, A).reshape(chi * 4, chi * 9)
    assert H.maxrel(orc.c2x2('LU', C, T1, T2, a), ref) < 1e-14


def test_multiplet_rule_edge_cases():
    # cut inside a degenerate pair -> drop the whole multiplet (custom_svd.py:70-95)
    s = torch.tensor([1.0, 0.5, 0.5, 0.1], dtype=torch.float64)
    assert orc.multiplet_chi(s, 2, 1e-8, 1e-14) == 0
    assert orc.multiplet_chi(s, 3, 1e-8, 1e-14) == 3
    # values below abs_tol count as zero
    s = torch.tensor([1.0, 1e-15, 1e-15, 0.0], dtype=torch.float64)
    assert orc.multiplet_chi(s, 2, 1e-8, 1e-14) == 0


# ------------------------------------------------------------------------------------------
# variants of the move and density matrices: reference outputs in tests/golden/variants_*.npz
# (oracle/gen_golden_variants.py) against the oracle, through the drop-in modules with the oracle as engine
# ------------------------------------------------------------------------------------------
@pytest.fixture()
def oracle_engine_everywhere(monkeypatch):
    from peps_torch_b200.ctm.generic import ctmrg, rdm
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v, rdm_c4v
    e = H.OracleEngine()
    for m in (ctmrg, rdm, ctmrg_c4v, rdm_c4v):
        monkeypatch.setattr(m, '_engine', lambda: e)
    return e


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128'])
def test_oracle_against_reference_variant_fixtures_generic(oracle_engine_everywhere, name):
    n = H.check_generic_variants(name, torch.device('cpu'), tol_move=1e-9, tol_rdm=1e-13, tol_rdm_spd=1e-13)
    assert n == 3 * 4 * 12 + 4 * 2 * 4


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_oracle_against_reference_variant_fixtures_c4v(oracle_engine_everywhere, name):
    assert H.check_c4v_variants(name, torch.device('cpu'), tol_C=1e-12, tol_T=1e-10, tol_rdm=1e-12, tol_rdm_spd=1e-12) == 12


C1_FINAL_ENERGY = -0.35003258049356745      # `ctmrg_j1j2_c4v.py --bond_dim 2 --chi 16 --seed 123 --j2 0.3` run here


def test_config1_script_known_answer_oracle():
    """BASELINE.json configs[0]: the reference script converges in four moves (rdm2x1 distance 7.2e-9 < ctm_conv_tol) and
    prints FINAL -0.35003258049356745; the oracle on the script-exact state (seed 123, family A) reproduces it."""
    a = orc.random_state_c4v(2, family='A')
    C, T = orc.init_env_c4v(a, 16)
    for _ in range(4):
        C, T = orc.ctm_move_c4v(a, C, T, 16)
    assert abs(orc.energy_j1j2_c4v(a, C, T, 1.0, 0.3) - C1_FINAL_ENERGY) < 1e-13


C2_FINAL_ENERGY = 0.6424192641900255    # `ctmrg_j1j2.py --tiling 4SITE --bond_dim 3 --chi 48 --seed 123 --j2 0.3`, three iterations


def test_config2_script_known_answer_oracle():
    """BASELINE.json configs[1] on the script-exact state (seed 123, family A): the reference script (moves of the
    reference, energy through rdm2x2 -- tests/launcher_probe.py, which has to supply rdm2x2 because the reference's own
    dispatch needs opt_einsum) converges in three iterations and prints FINAL 0.6424192641900255."""
    sites = orc.random_state_4site(3, family='A')
    C, T = orc.init_env(sites, orc.v2s_4site, 48)
    for _ in range(3):
        orc.ctm_iteration(sites, orc.v2s_4site, 2, 2, C, T, 48)
    assert abs(orc.energy_j1j2(sites, orc.v2s_4site, C, T, 1.0, 0.3) - C2_FINAL_ENERGY) < 1e-13
