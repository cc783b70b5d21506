"""Helper process of tests/test_ad_cpu.py: runs an unmodified reference OPTIMISATION script through
peps_torch_b200.run.enable() with the oracle standing in for libctmb (host-logic test of the autograd path: the moves are
built by peps_torch_b200/ad.py from the engine's einsum2 / truncated_svd / truncated_eig_sym).  With --plain the script runs
untouched (the reference's own moves), which gives the numbers to compare with."""
import os
import runpy
import sys

plain = sys.argv[1] in ('--plain', '--plain-legacy-rdm')
legacy = sys.argv[1] == '--plain-legacy-rdm'      # untouched moves; only rdm2x2 -> rdm2x2_legacy (opt_einsum is absent)
script = sys.argv[2] if plain else sys.argv[1]
rest = sys.argv[3:] if plain else sys.argv[2:]
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(repo, 'oracle'), os.path.join(repo, 'tests'), repo):
    sys.path.insert(0, p)
from peps_torch_b200 import run as launcher   # noqa: E402

root = launcher.find_reference_root(script)
for p in (os.path.dirname(os.path.abspath(script)), root):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True
if legacy:
    import importlib   # noqa: E402
    _rdm = importlib.import_module('ctm.generic.rdm')
    _legacy = _rdm.rdm2x2_legacy
    _rdm.rdm2x2 = lambda coord, state, env, open_sites=[0, 1, 2, 3], sym_pos_def=False, **kw: _legacy(coord, state, env, sym_pos_def=sym_pos_def)
    _rdm.rdm1x1, _rdm.rdm2x1, _rdm.rdm1x2 = _rdm.rdm1x1_dl, _rdm.rdm2x1_dl, _rdm.rdm1x2_dl      # the patch of SURVEY 8c, caveat 1
if not plain:
    import helpers as H   # noqa: E402
    from peps_torch_b200 import ad   # noqa: E402
    calls = {'ad_c4v': 0, 'ad_generic': 0}
    orig_c4v, orig_gen = ad.ctm_move_c4v, ad.ctm_move_generic

    def c4v(*a, **kw):
        calls['ad_c4v'] += 1
        return orig_c4v(*a, **kw)

    def gen(*a, **kw):
        calls['ad_generic'] += 1
        return orig_gen(*a, **kw)
    ad.ctm_move_c4v, ad.ctm_move_generic = c4v, gen
    launcher.enable(engine_factory=lambda: H.OracleEngine())
sys.argv = [script] + rest
try:
    runpy.run_path(script, run_name='__main__')
finally:
    if not plain:
        print('AD_CALLS', calls['ad_c4v'], calls['ad_generic'])
