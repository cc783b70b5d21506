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
