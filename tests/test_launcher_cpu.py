"""The launcher (peps_torch_b200/run.py) rebinds ctm_MOVE / ctm_MOVE_sl inside an unmodified
reference script.  Needs the reference tree, which exists only in the build container: skipped elsewhere."""
import os
import subprocess
import sys
import pytest

REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'examples', 'j1j2')),
                                reason='reference tree not present (GPU box)')


def _run(script, args, tmp_path):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='2')
    out = subprocess.run([sys.executable, os.path.join(HERE, 'launcher_probe.py'), os.path.join(REF, script)] + args,
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith('LAUNCHER_CALLS')][-1].split()
    return int(line[1]), int(line[2]), int(line[3]), out.stdout


def test_generic_script_calls_the_rebound_move(tmp_path):
    g, c, r, out = _run('examples/j1j2/ctmrg_j1j2.py', ['--tiling', '4SITE', '--bond_dim', '2', '--chi', '8', '--seed', '123',
                                                     '--j2', '0.3', '--CTMARGS_ctm_max_iter', '2'], tmp_path)
    assert g == 16 and c == 0          # 2 iterations x 2(lX+lY) moves of the 2x2 cell
    assert 'FINAL' in out and r >= 4   # the energy evaluation ran through the rebound rdm2x2 (one call per plaquette)


def test_c4v_script_calls_the_rebound_move(tmp_path):
    g, c, r, out = _run('examples/j1j2/ctmrg_j1j2_c4v.py', ['--bond_dim', '2', '--chi', '8', '--seed', '123', '--j2', '0.3',
                                                         '--CTMARGS_ctm_max_iter', '3'], tmp_path)
    assert g == 0 and c >= 1
    assert 'FINAL' in out and r >= 1   # energy_1x1_lowmem through the rebound rdm2x2_NN(N)_lowmem_sl


def test_kagome_script_calls_the_rebound_move(tmp_path):
    """BASELINE config 4: the kagome iPESS script drives ctm.generic.ctmrg.run on a 1x1 cell with a p = 8 on-site tensor
    (SURVEY 3.3); complex128 because the CLI's --jperm is complex (caveat 2)."""
    g, c, r, out = _run('examples/kagome/ctmrg_spin_half_kagome.py', ['--ansatz', 'IPESS', '--bond_dim', '2', '--chi', '8',
                                                                      '--seed', '123', '--CTMARGS_ctm_max_iter', '2',
                                                                      '--GLOBALARGS_dtype', 'complex128'], tmp_path)
    assert g == 8 and c == 0           # 2 iterations x 2(lX+lY) moves of the 1x1 cell
    assert 'spectrum(T)' in out        # the script ran to its end (observables, transfer-matrix spectra)


KAGOME_RVB_REF = """-0.3931221584692804, (-0.5896832690555696+0j), (-0.5896832063522717+0j), (7.59716523160245e-32+0j),
    (-4.331814810151939e-31+0j), (8.592175414886632e-32+0j), (3.07218410812194e-16+0j),
    (-3.674157727896386e-17+0j), (5.011080358949883e-16+0j), (-3.953569086037768e-16+0j),
    (6.82165674627862e-16+0j), (-8.641428147458597e-16+0j), (-1.0617693286257969e-16+0j),
    (-9.980184244909592e-16+0j), (-7.479642784634561e-17+0j), (-0.19656089027856907+0j),
    (-0.19656149813919332+0j), (-0.1965608806378064+0j), (-0.19656141466722352+0j),
    (-0.19656089010487604+0j), (-0.19656090158017214+0j)"""


def check_kagome_rvb_final(stdout):
    """The reference's own golden vector for the kagome RVB state (examples/kagome/ctmrg_spin_half_kagome.py:405-417,
    TestCtmrg_IPESS_D3_RVB: 21 numbers at 1e-6): energy per site, down / up triangle energies, magnetisations, bond SS."""
    from cmath import isclose
    final = [ln for ln in stdout.splitlines() if ln.startswith('FINAL')]
    assert final, stdout[-2000:]
    got = [complex(x) for x in final[-1][len('FINAL'):].split(',')]
    want = [complex(x) for x in KAGOME_RVB_REF.split(',')]
    assert len(got) >= len(want)
    for g, w in zip(got, want):
        assert isclose(g, w, rel_tol=1e-6, abs_tol=1e-6), (g, w)


def test_kagome_rvb_known_answer_through_the_launcher(tmp_path):
    """Unmodified kagome script on the reference's RVB test state (IPESS D=3, chi=18, complex128) with the moves, the kagome
    density matrices (peps_torch_b200/ctm/pess_kagome/rdm_kagome.py) and the transfer-operator spectra rebound -- the oracle
    standing in for libctmb -- must reproduce the reference's golden vector.  (The GPU suite runs the same with libctmb.)"""
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(HERE, 'launcher_probe_ad.py'),
                          os.path.join(REF, 'examples', 'kagome', 'ctmrg_spin_half_kagome.py'), '--ansatz', 'IPESS', '--instate',
                          os.path.join(REF, 'test-input', 'IPESS_KAGOME_D3_RVB.in'), '--bond_dim', '3', '--chi', '18', '--j1', '1.0',
                          '--GLOBALARGS_dtype', 'complex128', '--out_prefix', 'kg'],
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    check_kagome_rvb_final(out.stdout)


def _ss_lines(stdout):
    rows, take = [], False
    for ln in stdout.splitlines():
        if ln.startswith(('SS[(', 'SS r', 'DD r', 'DD_v r', 'spectrum(T)', 'spectrum(T2)')):
            take = True
            continue
        if take:
            f = ln.split()
            if len(f) >= 2 and f[0].isdigit():
                rows.append([complex(x) for x in f[1:]])
            elif ln.strip():
                take = False
    return rows


def test_c4v_correlation_functions_of_the_script_through_the_launcher(tmp_path):
    """SS and dimer-dimer correlation functions and the transfer-operator spectrum at the tail of ctmrg_j1j2_c4v.py
    (models/j1j2.py:826-925 -> corrf_c4v.corrf_1sO1sO / corrf_2sOH2sOH_E1 / corrf_2sOV2sOV_E2, transferops_c4v.get_Top_spec_c4v /
    get_Top2_spec_c4v) through the
    launcher with the oracle as engine against the script run untouched."""
    args = ['--bond_dim', '2', '--chi', '8', '--seed', '123', '--j2', '0.3', '--j3', '0.1', '--CTMARGS_ctm_max_iter', '6',
            '--corrf_r', '4', '--top_n', '3', '--corrf_dd_v', '--top2']
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='2')
    script = os.path.join(REF, 'examples', 'j1j2', 'ctmrg_j1j2_c4v.py')
    outs, finals = [], []
    for mode in (['--plain'], []):
        wd = tmp_path / ('a' if mode else 'b')
        os.makedirs(wd, exist_ok=True)
        out = subprocess.run([sys.executable, os.path.join(HERE, 'launcher_probe_ad.py')] + mode + [script] + args,
                             cwd=wd, env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(_ss_lines(out.stdout))
        finals.append([complex(x) for x in [ln for ln in out.stdout.splitlines() if ln.startswith('FINAL')][-1][6:].split(',')])
    # j3 != 0: the energy goes through rdm3x1_sl as well (models/j1j2.py:671-677)
    assert len(finals[0]) == len(finals[1]) and all(abs(x - y) < 1e-9 for x, y in zip(*finals)), finals
    want, got = outs
    assert len(want) == 4 + 4 + 4 + 3 + 3 and len(got) == len(want)     # SS, DD, DD_v rows, spectrum(T), spectrum(T2) rows
    for a, b in zip(got, want):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert abs(x - y) < 1e-8 * max(1.0, abs(y)), (got, want)


def test_correlation_functions_of_the_script_through_the_launcher(tmp_path):
    """The SS correlation functions printed at the tail of ctmrg_j1j2.py (models/j1j2.py:476-508 -> corrf_1sO1sO) through the
    launcher (get_edge / apply_TM_1sO with operators / apply_edge rebound, oracle standing in for libctmb) against the script
    run untouched."""
    args = ['--tiling', '4SITE', '--bond_dim', '2', '--chi', '8', '--seed', '123', '--j2', '0.3', '--CTMARGS_ctm_max_iter', '4',
            '--corrf_r', '5']
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='2')
    script = os.path.join(REF, 'examples', 'j1j2', 'ctmrg_j1j2.py')
    outs = []
    for mode in (['--plain-legacy-rdm'], []):
        for sub in ('a', 'b'):
            os.makedirs(tmp_path / sub, exist_ok=True)
        out = subprocess.run([sys.executable, os.path.join(HERE, 'launcher_probe_ad.py')] + mode + [script] + args,
                             cwd=tmp_path / ('a' if mode else 'b'), env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(_ss_lines(out.stdout))
    want, got = outs
    assert len(want) >= 10 and len(got) == len(want)      # 5 distances x 2 directions, then the spectrum(T) rows
    for a, b in zip(got, want):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert abs(x - y) < 1e-9 * max(1.0, abs(y)), (got, want)
