"""The drop-in claim on the GPU: UNMODIFIED example scripts of the reference, driven through the launcher
(python -m peps_torch_b200.run <script> ... --GLOBALARGS_device cuda:0), with libctmb as the CTM engine, must print the
FINAL energies the reference itself prints on CPU for BASELINE configs 1 and 2 (the numbers were produced
by running the unmodified scripts in the build container).

The scripts draw their random state with torch.rand on the device named by --GLOBALARGS_device, so `--seed 123` is a different
state on cuda:0 than on the CPU (measured: FINAL -0.3577 / 0.6470 instead of -0.3500 / 0.6424).  The CPU seed-123 states are
therefore fed through the scripts' own --instate option from tests/golden/config{1,2}_instate.json, written by the reference's
writer (oracle/gen_instates.py, which also checks that the unmodified script prints the identical FINAL line from the file).

Needs the staged copy of the reference under baseline/_ref/ (git-ignored; written by __graft_entry__.build() where the
reference tree exists, travels to the GPU box with the snapshot): skipped where it is absent."""
import os
import subprocess
import sys
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _final(script, args, tmp_path):
    if not os.path.isfile(os.path.join(REF, script)):
        pytest.skip('baseline/_ref (staged copy of the reference) not present')
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a CUDA device')
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    out = subprocess.run([sys.executable, '-m', 'peps_torch_b200.run', os.path.join(REF, script)] + args
                         + ['--GLOBALARGS_device', 'cuda:0'], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('FINAL')]
    assert lines, out.stdout[-2000:]
    return [float(x) for x in lines[-1][len('FINAL'):].split(',')], out.stdout


def test_config1_script_unmodified_on_gpu(tmp_path):
    """BASELINE configs[0]: examples/j1j2/ctmrg_j1j2_c4v.py --bond_dim 2 --chi 16 --seed 123 --j2 0.3 (converges at the fourth
    move by its own rdm2x1 criterion, ctmrg_j1j2_c4v.py:101-129) -> FINAL -0.35003258049356745, ..."""
    vals, out = _final('examples/j1j2/ctmrg_j1j2_c4v.py', ['--instate', os.path.join(GOLD, 'config1_instate.json'), '--chi', '16',
                                                           '--j2', '0.3'], tmp_path)
    assert abs(vals[0] - (-0.35003258049356745)) < 1e-10 * 0.35, vals[0]
    # the script's own convergence history: four moves, as on CPU
    iters = [ln for ln in out.splitlines() if ln[:1].isdigit() and ', ' in ln and len(ln.split(', ')) == 2]
    assert len(iters) == 4, iters


def test_config2_script_unmodified_on_gpu(tmp_path):
    """BASELINE configs[1]: examples/j1j2/ctmrg_j1j2.py --tiling 4SITE --bond_dim 3 --chi 48 --seed 123 --j2 0.3 -> FINAL
    0.6424192641900255, ... (energy per site from rdm2x2 of every plaquette, models/j1j2.py:223-247)."""
    vals, _ = _final('examples/j1j2/ctmrg_j1j2.py', ['--instate', os.path.join(GOLD, 'config2_instate.json'), '--tiling', '4SITE',
                                                   '--chi', '48', '--j2', '0.3'], tmp_path)
    assert abs(vals[0] - 0.6424192641900255) < 1e-10 * 0.64, vals[0]


def test_kagome_rvb_known_answer_unmodified_on_gpu(tmp_path):
    """The reference's own golden vector for the kagome family (BASELINE config 4's script; TestCtmrg_IPESS_D3_RVB,
    ctmrg_spin_half_kagome.py:360-417: IPESS D=3 chi=18 complex128 on test-input/IPESS_KAGOME_D3_RVB.in, 21 numbers at 1e-6):
    the unmodified script through the launcher on cuda:0 -- moves, kagome density matrices and transfer-operator mat-vecs on
    libctmb."""
    from test_launcher_cpu import check_kagome_rvb_final
    script = os.path.join(REF, 'examples', 'kagome', 'ctmrg_spin_half_kagome.py')
    instate = os.path.join(REF, 'test-input', 'IPESS_KAGOME_D3_RVB.in')
    if not (os.path.isfile(script) and os.path.isfile(instate)):
        pytest.skip('baseline/_ref (staged copy of the reference) not present')
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a CUDA device')
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    out = subprocess.run([sys.executable, '-m', 'peps_torch_b200.run', script, '--ansatz', 'IPESS', '--instate', instate,
                          '--bond_dim', '3', '--chi', '18', '--j1', '1.0', '--GLOBALARGS_dtype', 'complex128', '--out_prefix', 'kg',
                          '--GLOBALARGS_device', 'cuda:0'], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    check_kagome_rvb_final(out.stdout)


# the reference's own golden vectors for the generic J1-J2 script (TestCtmrg_States, examples/j1j2/ctmrg_j1j2.py:244-308, 1e-6)
J1J2_STATES = [
    (['--tiling', 'BIPARTITE', '--j3', '0.125', '--h_uni', '3.9', '0', '0'],
     'BIPARTITE_j2_0_j3_1250_h_39000_D_3_chi_32_seed_100_state.json',
     """-1.3896897615463615, 0.4884474386344192, 0.48844697363007333, 0.4884479036387651,
        -0.46200561021924863, 0.1585284270227813, 0.1585284270227813, -0.4620060817268178,
        -0.15852991836412875, -0.15852991836412875, 0.1751621217404098, 0.17516618332251627,
        0.17516347390256323, 0.17516132836311246"""),
    (['--tiling', '2SITE', '--j2', '0.55'],
     'gesdd-D2-chi50-j20.55-run0-iRND2x1_state.json',
     """-0.4434603770143078, 0.3184895704619597, 0.31842030538406385, 0.31855883553985553,
        -0.26397659399034457, 0.17806697814624955, 0.1780669781462514, 0.26446699770693394,
        -0.17758642635176156, -0.1775864263517598"""),
]


@pytest.mark.parametrize('case', [0, 1])
def test_j1j2_states_known_answers_unmodified_on_gpu(tmp_path, case):
    """TestCtmrg_States of the reference: BIPARTITE D=3 chi=32 with j3 and a field (14 numbers) and 2SITE D=2 chi=32 at
    j2 = 0.55 (10 numbers: energy, magnetisations, bond correlators), unmodified script through the launcher on cuda:0."""
    from cmath import isclose
    args, instate, ref = J1J2_STATES[case]
    script = os.path.join(REF, 'examples', 'j1j2', 'ctmrg_j1j2.py')
    f = os.path.join(REF, 'test-input', instate)
    if not (os.path.isfile(script) and os.path.isfile(f)):
        pytest.skip('baseline/_ref (staged copy of the reference) not present')
    vals, _ = _final('examples/j1j2/ctmrg_j1j2.py', args + ['--instate', f, '--chi', '32', '--CTMARGS_ctm_max_iter', '100'], tmp_path)
    want = [float(x) for x in ref.split(',')]
    assert len(vals) >= len(want)
    for g, w in zip(vals, want):
        assert isclose(g, w, rel_tol=1e-6, abs_tol=1e-6), (vals, want)
