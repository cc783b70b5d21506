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
