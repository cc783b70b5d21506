"""Write the script-exact random states of BASELINE configs 1 and 2 as instate files (tests/golden/config{1,2}_instate.json).

Run in the BUILD container only (needs /root/reference):   python oracle/gen_instates.py

The example scripts draw their random state with torch.rand ON THE DEVICE named by --GLOBALARGS_device
(examples/j1j2/ctmrg_j1j2_c4v.py:61-66, ctmrg_j1j2.py:80-96), so `--seed 123` gives a different state on cuda:0 than on
the CPU and the FINAL energies the reference prints on CPU cannot be compared with a GPU run of the same command line.
The files written here hold the CPU seed-123 states (the ones oracle/ctm_oracle.py: random_state_c4v / random_state_4site
family 'A' reproduce), serialised by the reference's own writer; tests/test_gpu_launcher.py feeds them to the UNMODIFIED
scripts through --instate.  The script then prints, on CPU, exactly the FINAL line of the seed-123 run (checked below)."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
GOLD = os.path.join(HERE, '..', 'tests', 'golden')
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
os.chdir(tempfile.mkdtemp())

import torch                                                  # noqa: E402
import ctm_oracle as orc                                      # noqa: E402
import config as cfg                                          # noqa: E402
from ipeps.ipeps import IPEPS                                 # noqa: E402
from ipeps.ipeps_c4v import IPEPS_C4V                         # noqa: E402

cfg.global_args.torch_dtype, cfg.global_args.dtype, cfg.global_args.device = torch.float64, 'float64', 'cpu'
f1 = os.path.abspath(os.path.join(GOLD, 'config1_instate.json'))
f2 = os.path.abspath(os.path.join(GOLD, 'config2_instate.json'))
IPEPS_C4V(orc.random_state_c4v(2, family='A')).write_to_file(f1)
sites = orc.random_state_4site(3, family='A')
IPEPS(sites, vertexToSite=orc.v2s_4site, lX=2, lY=2).write_to_file(f2)


def final(script, args):
    out = subprocess.run([sys.executable, os.path.join(REF, script)] + args, capture_output=True, text=True,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1'))
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.startswith('FINAL')][-1]


a = final('examples/j1j2/ctmrg_j1j2_c4v.py', ['--bond_dim', '2', '--chi', '16', '--seed', '123', '--j2', '0.3'])
b = final('examples/j1j2/ctmrg_j1j2_c4v.py', ['--instate', f1, '--chi', '16', '--j2', '0.3'])
print(a); print(b)
assert a == b, 'the instate file does not reproduce the seed-123 run'

# the state of the reference's optimisation test TestOpt4SITE (examples/j1j2/optim_j1j2.py:371-440: 4SITE, D = 2, seed 123)
f3 = os.path.abspath(os.path.join(GOLD, 'opt4site_instate.json'))
IPEPS(orc.random_state_4site(2, family='A'), vertexToSite=orc.v2s_4site, lX=2, lY=2).write_to_file(f3)
print('written', f3)
