"""Dry run of the GPU parity suite on CPU: the BODIES of the tests in tests/test_gpu_parity.py are executed with the oracle
standing in for libctmb (helpers.OracleEngine patched into the drop-in modules, device = cpu).  This checks the test
code itself -- fixtures, keys, shapes, tolerances that must hold between the oracle and the reference outputs -- before any
GPU time is spent on it; it says nothing about libctmb.  Tests that drive CUDA-only machinery (debug switches of the
library, large range-finder sketches, error paths of the C ABI) are not part of the dry run."""
import inspect
import pytest
import torch
import helpers as H
import test_gpu_parity as G
import test_gpu_parity_sizes as GS

DRY = ['test_pieces_against_reference_fixtures', 'test_run_energy_and_spectra_against_reference',
       'test_c4v_against_reference_fixtures', 'test_known_answer_rvb_c4v_on_gpu', 'test_known_answer_j1j2_2site_on_gpu',
       'test_projector_method_4x2_against_oracle', 'test_double_layer_sites_against_oracle',
       'test_2norm_normalisation_against_oracle', 'test_run_overlap_against_oracle',
       'test_run_force_dl_and_warmup_match_plain_run', 'test_c4v_double_layer_move_and_2norm',
       'test_sym_pos_def_matrix_against_oracle', 'test_rdm2x2_and_energy_against_oracle',
       'test_c4v_rdms_and_energy_against_reference_fixture', 'test_small_rdms_against_oracle',
       'test_c4v_small_rdms_against_oracle', 'test_variants_against_reference_fixtures_generic',
       'test_variants_against_reference_fixtures_c4v', 'test_config1_script_known_answer_on_gpu',
       'test_config2_script_known_answer_on_gpu', 'test_rectangular_cell_with_direction_dependent_bonds',
       # tests/test_gpu_parity_sizes.py
       'test_config3_size_c4v_complex_against_live_oracle', 'test_config4_size_kagome_against_live_oracle',
       'test_shard_entry_points_on_one_gpu', 'test_halves_against_reference_fixtures', 'test_clustered_spectrum_orthogonality']


def _find(name):
    return getattr(G, name, None) or getattr(GS, name)


def _cases():
    out = []
    for name in DRY:
        f = _find(name)
        marks = [m for m in getattr(f, 'pytestmark', []) if m.name == 'parametrize']
        if not marks:
            out.append(pytest.param(name, {}, id=name))
            continue
        argnames = [a.strip() for a in marks[0].args[0].split(',')]
        for vals in marks[0].args[1]:
            vals = vals if isinstance(vals, (tuple, list)) and len(argnames) > 1 else (vals,)
            kw = dict(zip(argnames, vals))
            out.append(pytest.param(name, kw, id=f'{name}[{"-".join(str(v) for v in vals)}]'))
    return out


@pytest.fixture()
def oracle_everywhere(monkeypatch):
    from peps_torch_b200.ctm.generic import ctmrg, rdm
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v, rdm_c4v
    e = H.OracleEngine()
    for m in (ctmrg, rdm, ctmrg_c4v, rdm_c4v):
        monkeypatch.setattr(m, '_engine', lambda: e)
    return e


@pytest.mark.parametrize('name,kw', _cases())
def test_gpu_test_body_runs_with_the_oracle_as_engine(oracle_everywhere, name, kw):
    f = _find(name)
    params = inspect.signature(f).parameters
    args = dict(kw)
    if 'eng' in params:
        args['eng'] = oracle_everywhere
    if 'dev' in params:
        args['dev'] = torch.device('cpu')
    f(**args)
