"""Kagome density matrices (BASELINE config 4's energy): peps_torch_b200/ctm/pess_kagome/rdm_kagome.py against
the unmodified reference (ctm/pess_kagome/rdm_kagome.py: trace1x1_dn_kagome, rdm2x2_dn_triangle_with_operator,
rdm2x2_up_triangle_open) and the energies of models/spin_half_kagome.py computed from them, with the oracle standing in for
libctmb.  Needs the reference tree (build container only); the GPU suite compares libctmb with the oracle-engine result."""
import copy
import os
import sys
import pytest
import torch
import helpers as H

REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ctm', 'pess_kagome')), reason='reference tree not present')


@pytest.fixture()
def ref(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import config as cfg
    saved = copy.deepcopy(cfg.global_args.__dict__)
    cfg.global_args.dtype, cfg.global_args.torch_dtype, cfg.global_args.device = 'float64', torch.float64, 'cpu'
    try:
        yield cfg
    finally:
        cfg.global_args.__dict__.clear()
        cfg.global_args.__dict__.update(saved)
        os.chdir(cwd)
        sys.path.remove(REF)


def kagome_fixture():
    z, meta = H.load_golden('kagome_1site_D2_chi8_A')
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'final_' if any(k.startswith('final_') for k in z.files) else 'mid_')
    return sites, v2s, lX, lY, C, T, meta['chi']


def test_kagome_rdms_and_energies_match_reference(ref, monkeypatch):
    from ipeps.ipeps_kagome import IPEPS_KAGOME
    from ctm.generic.env import ENV as RE
    from ctm.pess_kagome import rdm_kagome as rk
    from models import spin_half_kagome
    from peps_torch_b200.ctm.pess_kagome import rdm_kagome as ok
    eng = H.OracleEngine()
    monkeypatch.setattr(ok, '_engine', lambda: eng)
    sites, v2s, lX, lY, C, T, chi = kagome_fixture()
    rs = IPEPS_KAGOME(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
    re = RE(chi, rs)
    re.C, re.T = dict(C), dict(T)
    st, env = H.State(sites, v2s, lX, lY), H.Env(chi, dict(C), dict(T))
    g = torch.Generator().manual_seed(3)
    for op in (torch.randn(8, 8, dtype=torch.float64, generator=g), torch.randn([2] * 6, dtype=torch.float64, generator=g)):
        a, b = rk.trace1x1_dn_kagome((0, 0), rs, re, op), ok.trace1x1_dn_kagome((0, 0), st, env, op)
        assert abs(float(a) - float(b)) < 1e-12 * abs(float(a))
        (va, na), (vb, nb) = rk.rdm2x2_dn_triangle_with_operator((0, 0), rs, re, op), ok.rdm2x2_dn_triangle_with_operator((0, 0), st, env, op)
        assert abs(float(va) - float(vb)) < 1e-12 and abs(float(na) - float(nb)) < 1e-12 * abs(float(na))
    for spd in (False, True):
        want, got = rk.rdm2x2_up_triangle_open((0, 0), rs, re, sym_pos_def=spd), ok.rdm2x2_up_triangle_open((0, 0), st, env, sym_pos_def=spd)
        assert got.shape == want.shape == (2,) * 6
        assert float((want - got).abs().max()) < 1e-13
    # the energies of the model, with the reference's rdm_kagome functions rebound to ours (what the launcher does)
    model = spin_half_kagome.S_HALF_KAGOME(j1=1.0, jperm=0.)
    e_dn_ref, _ = model.energy_triangle_dn(rs, re)
    e_up_ref, _ = model.energy_triangle_up(rs, re)
    for name in ('trace1x1_dn_kagome', 'rdm2x2_dn_triangle_with_operator', 'rdm2x2_up_triangle_open'):
        monkeypatch.setattr(rk, name, getattr(ok, name))
    e_dn, _ = model.energy_triangle_dn(st, env)
    e_up, _ = model.energy_triangle_up(st, env)
    for c in e_dn_ref:
        assert abs(float(e_dn[c]) - float(e_dn_ref[c])) < 1e-12 and abs(float(e_up[c]) - float(e_up_ref[c])) < 1e-12
