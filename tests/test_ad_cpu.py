"""Reverse-mode AD through the CTM move, CPU part: the adjoints and chains of peps_torch_b200/ad.py against
gradients written by the UNMODIFIED reference (tests/golden/grad_*.npz, oracle/gen_golden_grad.py: loss.backward() through
ctmrg_c4v.run / ctmrg.run, i.e. SYMEIG.backward linalg/eig_sym.py:56-78 and SVDGESDD.backward linalg/svd_gesdd.py:209-328),
with the oracle standing in for libctmb as the engine behind einsum2 / truncated_svd / truncated_eig_sym.  The GPU suite
(tests/test_gpu_ad.py) runs the same checker with libctmb."""
import os
import subprocess
import sys
import pytest
import helpers as H

GRAD = ['grad_c4v_D2_chi16', 'grad_c4v_D2_chi16_c128', 'grad_c4v_D2_chi16_ckpt', 'grad_generic_4site_D2_chi8',
        'grad_generic_4site_D2_chi6_c128']
REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('name', GRAD)
def test_ad_move_matches_reference_gradient(name):
    de, dg, scale = H.check_grad_fixture(name, H.OracleEngine(), 'cpu', through_api=False)
    assert de < 1e-12 and dg < 1e-10 * max(1.0, scale), (name, de, dg, scale)


@pytest.mark.parametrize('name', ['grad_c4v_D2_chi16', 'grad_generic_4site_D2_chi8'])
def test_drop_in_move_dispatches_to_the_ad_path(name, monkeypatch):
    """ctm_MOVE / ctm_MOVE_sl called with tensors that require grad record an autograd graph (round 1 returned detached
    tensors); without grad they take the fused forward-only call."""
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    eng = H.OracleEngine()
    monkeypatch.setattr(ctmrg, '_engine', lambda: eng)
    monkeypatch.setattr(ctmrg_c4v, '_engine', lambda: eng)
    de, dg, scale = H.check_grad_fixture(name, eng, 'cpu', through_api=True)
    assert de < 1e-12 and dg < 1e-10 * max(1.0, scale), (name, de, dg, scale)
    assert not eng.calls                     # the fused (forward-only) entry points were never used


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'examples', 'j1j2')), reason='reference tree not present (GPU box)')
def test_optim_script_through_the_launcher(tmp_path):
    """examples/j1j2/optim_j1j2_c4v.py, unmodified, three L-BFGS steps: through the launcher (moves from peps_torch_b200/ad.py)
    it must print the energies it prints when run untouched."""
    args = ['--bond_dim', '2', '--chi', '8', '--seed', '123', '--j2', '0.3', '--opt_max_iter', '3',
            '--CTMARGS_ctm_max_iter', '6', '--out_prefix', 'adtest']
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='2')
    script = os.path.join(REF, 'examples', 'j1j2', 'optim_j1j2_c4v.py')
    outs = []
    for pre in ([], ['--plain']):
        d = tmp_path / ('plain' if pre else 'launcher')
        d.mkdir()
        out = subprocess.run([sys.executable, os.path.join(HERE, 'launcher_probe_ad.py')] + pre + [script] + args,
                             cwd=d, env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(out.stdout)
    calls = [ln for ln in outs[0].splitlines() if ln.startswith('AD_CALLS')][-1].split()
    assert int(calls[1]) >= 6          # every move of the differentiated CTM runs went through ad.ctm_move_c4v

    def energies(text):
        rows = [ln.split(', ') for ln in text.splitlines() if ln[:1].isdigit() or ln[:2] == '-1']
        return [float(r[1]) for r in rows if len(r) > 3]
    e_l, e_p = energies(outs[0]), energies(outs[1])
    assert len(e_l) == len(e_p) >= 4
    assert max(abs(a - b) for a, b in zip(e_l, e_p)) < 1e-10, (e_l, e_p)
    assert e_l[-1] < e_l[0] - 1e-4     # the optimiser moved downhill


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'examples', 'j1j2')), reason='reference tree not present (GPU box)')
def test_generic_optim_script_through_the_launcher(tmp_path):
    """examples/j1j2/optim_j1j2.py (4SITE), unmodified, two L-BFGS steps: through the launcher every differentiated move goes
    through ad.ctm_move_generic and the energies equal those of the untouched script (whose density matrices are switched to
    the reference's own opt_einsum-free variants, SURVEY 8c caveat 1 -- opt_einsum is not installed here)."""
    args = ['--tiling', '4SITE', '--bond_dim', '2', '--chi', '4', '--seed', '123', '--j2', '0.3', '--opt_max_iter', '2',
            '--CTMARGS_ctm_max_iter', '2', '--out_prefix', 'adg']
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='2')
    script = os.path.join(REF, 'examples', 'j1j2', 'optim_j1j2.py')
    outs = []
    for pre in ([], ['--plain-legacy-rdm']):
        d = tmp_path / ('plain' if pre else 'launcher')
        d.mkdir()
        out = subprocess.run([sys.executable, os.path.join(HERE, 'launcher_probe_ad.py')] + pre + [script] + args,
                             cwd=d, env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(out.stdout)
    calls = [ln for ln in outs[0].splitlines() if ln.startswith('AD_CALLS')][-1].split()
    assert int(calls[2]) >= 16         # 8 moves per CTM iteration, every differentiated iteration through ad.ctm_move_generic

    def energies(text):
        rows = [ln.split(', ') for ln in text.splitlines() if ln[:1].isdigit() or ln[:2] == '-1']
        return [float(r[1]) for r in rows if len(r) > 3]
    e_l, e_p = energies(outs[0]), energies(outs[1])
    assert len(e_l) == len(e_p) >= 3
    assert max(abs(a - b) for a, b in zip(e_l, e_p)) < 1e-10, (e_l, e_p)
    assert e_l[-1] < e_l[0] - 1e-3


@pytest.mark.parametrize('dt', ['float64', 'complex128'])
@pytest.mark.parametrize('shape', [(12, 12), (21, 21)])
def test_full_svd_adjoint_against_torch_autograd(dt, shape):
    """_SvdFull (the reference's regularised SVDGESDD.backward, linalg/svd_gesdd.py:209-328) against torch's own SVD autograd on
    a gauge-invariant function of (U, S, V); the regularisation 1e-12 is invisible on a well-separated spectrum.  Square
    matrices: the projector matrix M = R^T Rt of the move always is."""
    import torch
    from peps_torch_b200 import ad
    dtype = getattr(torch, dt)
    g = torch.Generator().manual_seed(7)
    A0 = torch.randn(shape, dtype=dtype, generator=g)
    Wm = torch.randn(shape, dtype=dtype, generator=g)
    k = min(shape)

    def f(U, S, V):
        # gauge invariant: depends on U S^2 V^H, U diag(w) U^H and S only
        w = torch.linspace(1.0, 2.0, k, dtype=S.dtype)
        P = (U * w.to(U.dtype)) @ U.conj().t()
        return ((U * (S ** 2).to(U.dtype)) @ V.conj().t() * Wm.conj()).sum().real + P.abs().pow(2).sum() + (S ** 3).sum()
    A1 = A0.clone().requires_grad_(True)
    U, S, V = ad._SvdFull.apply(H.OracleEngine(), A1, 1.0e-12)
    f(U[:, :k], S, V[:, :k]).backward()
    A2 = A0.clone().requires_grad_(True)
    U2, S2, Vh2 = torch.linalg.svd(A2, full_matrices=False)
    f(U2, S2, Vh2.conj().t()).backward()
    assert float((A1.grad - A2.grad).abs().max()) < 1e-9 * float(A2.grad.abs().max())


@pytest.mark.parametrize('dt', ['float64', 'complex128'])
def test_full_eig_adjoint_against_torch_autograd(dt):
    """_EigSymFull (SYMEIG.backward, linalg/eig_sym.py:56-78) against torch.linalg.eigh's autograd."""
    import torch
    from peps_torch_b200 import ad
    dtype = getattr(torch, dt)
    g = torch.Generator().manual_seed(9)
    X = torch.randn(14, 14, dtype=dtype, generator=g)
    A0 = 0.5 * (X + X.conj().t())
    w = torch.linspace(1.0, 2.0, 14, dtype=torch.float64)

    def f(D, U):
        # eigenvalues come sorted by magnitude from ours, ascending from torch: use a permutation-invariant, gauge-invariant function
        return (D ** 3).sum() + ((U * torch.tanh(D).to(U.dtype)) @ U.conj().t()).abs().pow(2).sum() + ((U * D.to(U.dtype)) @ U.conj().t() * A0.conj()).sum().real
    A1 = A0.clone().requires_grad_(True)
    D, U = ad._EigSymFull.apply(H.OracleEngine(), 0.5 * (A1 + A1.conj().t()), 1.0e-12)
    f(D, U).backward()
    A2 = A0.clone().requires_grad_(True)
    D2, U2 = torch.linalg.eigh(0.5 * (A2 + A2.conj().t()))
    f(D2, U2).backward()
    assert float((A1.grad - A2.grad).abs().max()) < 1e-9 * float(A2.grad.abs().max())
    del w
