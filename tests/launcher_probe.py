"""Helper process of tests/test_launcher_cpu.py: runs an unmodified reference script through
peps_torch_b200.run.enable() with a RECORDING engine (host-logic test; the recording engine forwards
to the reference's own move so that the script's output stays meaningful)."""
import os
import runpy
import sys

script = sys.argv[1]
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(repo, 'oracle'), os.path.join(repo, 'tests'), repo):
    sys.path.insert(0, p)
import helpers as H   # noqa: E402
from peps_torch_b200 import run as launcher   # noqa: E402

root = launcher.find_reference_root(script)
for p in (os.path.dirname(os.path.abspath(script)), root):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True
import importlib   # noqa: E402
ref = importlib.import_module('ctm.generic.ctmrg')
ref_c4v = importlib.import_module('ctm.one_site_c4v.ctmrg_c4v')
orig_move, orig_move_sl = ref.ctm_MOVE, ref_c4v.ctm_MOVE_sl
calls = {'generic': 0, 'c4v': 0, 'rdm': 0}


class RecordingEngine(H.OracleEngine):          # density matrices: the oracle (host-logic test, no GPU here)
    device = 'cpu'

    def move_generic(self, direction, state, env, **opt):
        calls['generic'] += 1
        orig_move(direction, state, env)

    def move_c4v(self, a, C, T, chi, **opt):
        raise RuntimeError('not used: ctm_MOVE_sl is recorded at the module level below')

    def rdm2x2(self, *a, **kw):
        calls['rdm'] += 1
        return super().rdm2x2(*a, **kw)

    def rdm2x2_sites(self, *a, **kw):
        calls['rdm'] += 1
        return super().rdm2x2_sites(*a, **kw)


launcher.enable(engine_factory=lambda: RecordingEngine())
# C4v: the drop-in unpacks tensors for the engine; record at the function level instead
new_sl = ref_c4v.ctm_MOVE_sl


def counted_sl(a, env, *args, **kw):
    calls['c4v'] += 1
    assert new_sl.__module__ == 'peps_torch_b200.run'
    return orig_move_sl(a, env, *args, **kw)


ref_c4v.ctm_MOVE_sl = counted_sl
sys.argv = [script] + sys.argv[2:]
try:
    runpy.run_path(script, run_name='__main__')
finally:
    print('LAUNCHER_CALLS', calls['generic'], calls['c4v'], calls['rdm'])
