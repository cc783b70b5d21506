"""State I/O: read_ipeps / write_ipeps of peps_torch_b200/ipeps.py against the reference's JSON format
(ipeps/ipeps.py:339-441, 501-535; ipeps/tensor_io.py).  The files under tests/golden/config{1,2}_instate.json were written by
the reference's own writer (oracle/gen_instates.py)."""
import os
import sys
import pytest
import torch
from peps_torch_b200.ipeps import read_ipeps, write_ipeps

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')


@pytest.mark.parametrize('fmt', ['legacy', '1D'])
@pytest.mark.parametrize('name', ['config1_instate.json', 'config2_instate.json'])
def test_round_trip(tmp_path, name, fmt):
    st = read_ipeps(os.path.join(GOLD, name))
    f = tmp_path / 'out.json'
    write_ipeps(st, f, tensor_io_format=fmt)
    st2 = read_ipeps(f)
    assert (st2.lX, st2.lY) == (st.lX, st.lY) and list(st2.sites) == list(st.sites)
    for c in st.sites:
        assert torch.equal(st.sites[c], st2.sites[c])                   # repr() of a double round-trips exactly
    for c in [(0, 0), (3, 5), (-1, -2), (2, 1)]:
        assert st.vertexToSite(c) == st2.vertexToSite(c)
    # another order of the auxiliary legs in the file is undone on reading (aux_ind_seq travels with the file)
    write_ipeps(st, f, aux_seq=(1, 0, 3, 2), tensor_io_format=fmt)
    st3 = read_ipeps(f)
    for c in st.sites:
        assert torch.equal(st.sites[c], st3.sites[c])
    stc = read_ipeps(os.path.join(GOLD, name), dtype=torch.complex128)
    assert all(t.dtype == torch.complex128 for t in stc.sites.values())


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'ipeps', 'ipeps.py')), reason='reference tree not present (GPU box)')
def test_against_the_reference_reader_and_writer(tmp_path):
    import subprocess
    code = f'''
import sys, os
sys.path[:0] = [{os.path.dirname(HERE)!r}, {REF!r}]
os.chdir({str(tmp_path)!r})
import torch
import config as cfg
from ipeps.ipeps import read_ipeps as rref, write_ipeps as wref
from peps_torch_b200.ipeps import read_ipeps, write_ipeps
src = {os.path.join(GOLD, 'config2_instate.json')!r}
ours, ref = read_ipeps(src), rref(src)
assert all(torch.equal(ours.sites[c], ref.sites[c]) for c in ref.sites)
assert all(ours.vertexToSite(c) == ref.vertexToSite(c) for c in [(0, 0), (3, 5), (-1, -2), (2, 1)])
write_ipeps(ours, 'a.json'); r2 = rref('a.json')
assert all(torch.equal(ours.sites[c], r2.sites[c]) for c in ref.sites)
write_ipeps(ours, 'b.json', tensor_io_format='1D'); r3 = rref('b.json')
assert all(torch.equal(ours.sites[c], r3.sites[c]) for c in ref.sites)
wref(ref, 'c.json'); o2 = read_ipeps('c.json')
assert all(torch.equal(o2.sites[c], ref.sites[c]) for c in ref.sites)
print('OK')
'''
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1'))
    assert out.returncode == 0 and 'OK' in out.stdout, out.stderr[-2000:]
