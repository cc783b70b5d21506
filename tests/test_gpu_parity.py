"""GPU parity suite (pytest -m gpu, B200 box): the CUDA path through the C ABI against the
oracle and the committed reference fixtures.  Tolerances: deterministic contractions 1e-13;
quantities behind the truncated SVD/EVD are compared through gauge invariants at
max(1e-10, measured reference-vs-reference noise floor);
energy 1e-10 relative (north star)."""
from collections import OrderedDict
import numpy as np
import pytest
import torch
import ctm_oracle as orc
import helpers as H

pytestmark = pytest.mark.gpu

GENERIC = ['generic_4site_D2_chi8_A', 'generic_4site_D2_chi8_B', 'generic_4site_D3_chi12_B',
           'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A']
C4V = ['c4v_D2_chi8_A', 'c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128']


@pytest.fixture(scope='module')
def eng():
    from peps_torch_b200.engine import CtmEngine
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a CUDA device')
    return CtmEngine()


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


def cpu(d):
    return {k: v.cpu() for k, v in d.items()}


@pytest.mark.parametrize('dt', [torch.float64, torch.complex128])
def test_einsum2_shapes_and_conj(eng, dev, dt):
    torch.manual_seed(1)
    cases = [('ab,buc->auc', (7, 5), (5, 3, 11)), ('auc,ael->ucel', (13, 4, 9), (13, 9, 4)),
             ('ik,kj->ij', (130, 70), (70, 150)), ('ki,kj->ij', (200, 129), (200, 65)),
             ('pqcert,sprfg->qcetsfg', (3, 3, 5, 5, 3, 3), (2, 3, 3, 3, 3)),
             ('ik,kj->ji', (1, 1), (1, 1)), ('ik,kj->ij', (257, 33), (33, 31)), ('ik,kj->ij', (300, 500), (500, 260)),
             # long reductions on few tiles: split-K path (slices + deterministic reduction kernel)
             ('ik,kj->ij', (40, 5000), (5000, 24)), ('ki,kj->ji', (3001, 8), (3001, 70)), ('ik,jk->ij', (130, 2048), (9, 2048))]
    for spec, sa, sb in cases:
        A = torch.randn(sa, dtype=dt, device=dev)
        B = torch.randn(sb, dtype=dt, device=dev)
        for ca, cb in ((False, False), (True, False), (False, True)):
            out = eng.einsum2(spec, A, B, conjA=ca, conjB=cb)
            ref = torch.einsum(spec, A.conj() if ca else A, B.conj() if cb else B)
            assert H.maxrel(out.cpu(), ref.cpu()) < 1e-13, (spec, ca, cb)


@pytest.mark.parametrize('name', GENERIC)
def test_pieces_against_reference_fixtures(eng, dev, name):
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'mid_')
    coord = list(sites.keys())[-1]
    for kind in orc.CORNERS:
        kc, k1, k2, _ = orc.CORNERS[kind]
        out = eng.c2x2(kind, C[(coord, kc)].to(dev), T[(coord, k1)].to(dev), T[(coord, k2)].to(dev),
                       sites[coord].to(dev), chi)
        assert H.maxrel(out.cpu(), torch.from_numpy(z['c2x2_' + kind])) < 1e-13
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    for d in orc.DIRECTIONS:
        tg = f'{d[0]}_{d[1]}'
        R, Rt = torch.from_numpy(z[f'halves_{tg}_R']), torch.from_numpy(z[f'halves_{tg}_Rt'])
        Pr, Ptr = torch.from_numpy(z[f'proj_{tg}_P']), torch.from_numpy(z[f'proj_{tg}_Pt'])
        M = R.t() @ Rt
        U, S, V = eng.truncated_svd(M.to(dev), chi)
        Sr = torch.linalg.svdvals(M)
        # the fixture states are (numerically) rank deficient: compare what the projector uses
        keep = Sr[:chi] / Sr[0] > 1e-8
        assert float((S.cpu()[keep] - Sr[:chi][keep]).abs().max() / Sr[0]) < 1e-13
        P, Pt, S2 = eng.projectors(R.to(dev), Rt.to(dev), chi)
        assert H.maxrel((P @ Pt.t()).cpu(), Pr @ Ptr.t()) < 1e-9
        assert H.maxrel(P.abs().cpu(), Pr.abs()) < 1e-8
        env = H.Env(chi, H.to_dev(C, dev), H.to_dev(T, dev))
        before = {('C',) + k: (v, v.clone()) for k, v in env.C.items()}
        before.update({('T',) + k: (v, v.clone()) for k, v in env.T.items()})
        before.update({('a', c): (v, v.clone()) for c, v in st.sites.items()})
        eng.move_generic(d, st, env)
        Cg, Tg = H.golden_env(z, f'move_{tg}_')
        assert H.env_abs_diff(env.C, env.T, Cg, Tg) < 1e-8
        # inputs are never modified (ctmrg.py:317-319: dict entries are REPLACED by fresh tensors): every input tensor
        # still holds, bit for bit, what it held before the move, and the written entries are new objects
        for k, (t, copy) in before.items():
            assert torch.equal(t, copy), k
        kC1, kC2, kT = orc.ABSORB[d]['out']
        for k, v in list(env.C.items()) + list(env.T.items()):
            if k[1] in (kC1, kC2, kT):
                assert all(v.data_ptr() != t.data_ptr() for t, _ in before.values())


@pytest.mark.parametrize('name', [n for n in GENERIC if 'kagome' not in n])
def test_run_energy_and_spectra_against_reference(eng, dev, name):
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.config import CTMARGS
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'init_')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
    args = CTMARGS()
    args.ctm_max_iter = meta['n_iter'] + 1          # fixture: 1 + n_iter reference iterations
    env, hist, t_ctm, t_obs = ctmrg.run(st, env, ctm_args=args)
    Cf, Tf = H.golden_env(z, 'final_')
    assert H.spectra_diff(env.C, Cf) < 1e-9
    assert H.env_abs_diff(env.C, env.T, Cf, Tf) < 1.2e-8        # max(1e-10, 3 x 4e-9), SURVEY 8c
    e = orc.energy_j1j2(sites, v2s, cpu(env.C), cpu(env.T), 1.0, meta['j2'])
    e_ref = float(z['energy'][0])
    assert abs(e - e_ref) <= 1e-10 * abs(e_ref) + 1e-13
    assert t_ctm > 0


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128'])
def test_single_precision_move_within_1e4(eng, dev, name):
    """float32 / complex64 (ENV takes its dtype from the state, ctm/generic/env.py:78-83): north-star gate 1e-4 against the
    reference's float64 result on the same (rounded) inputs; outputs come back in the input precision."""
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    low = torch.complex64 if next(iter(sites.values())).is_complex() else torch.float32
    C, T = H.golden_env(z, 'mid_')
    st = H.State({c: t.to(low).to(dev) for c, t in sites.items()}, v2s, lX, lY)
    for d in orc.DIRECTIONS:
        env = H.Env(chi, {k: v.to(low).to(dev) for k, v in C.items()}, {k: v.to(low).to(dev) for k, v in T.items()})
        eng.move_generic(d, st, env)
        assert all(v.dtype == low for v in list(env.C.values()) + list(env.T.values()))
        Cg, Tg = H.golden_env(z, f'move_{d[0]}_{d[1]}_')
        assert H.env_abs_diff({k: v.double() if low == torch.float32 else v.to(torch.complex128) for k, v in env.C.items()},
                              {k: v.double() if low == torch.float32 else v.to(torch.complex128) for k, v in env.T.items()},
                              Cg, Tg) < 1e-4


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_single_precision_c4v_move_within_1e4(eng, dev, name):
    z, meta = H.load_golden(name)
    chi = meta['chi']
    a64 = torch.from_numpy(z['site'])
    low = torch.complex64 if a64.is_complex() else torch.float32
    nC, nT, _ = eng.move_c4v(a64.to(low).to(dev), torch.from_numpy(z['mid_C']).to(low).to(dev),
                             torch.from_numpy(z['mid_T']).to(low).to(dev), chi)
    assert nC.dtype == low and nT.dtype == low
    assert H.maxrel(nC.cpu().to(a64.dtype), torch.from_numpy(z['mid_nC'])) < 1e-4
    assert H.maxrel(nT.abs().cpu().double(), torch.from_numpy(z['mid_nT']).abs()) < 1e-4


@pytest.mark.parametrize('name', C4V)
def test_c4v_against_reference_fixtures(eng, dev, name):
    z, meta = H.load_golden(name)
    chi = meta['chi']
    a = torch.from_numpy(z['site']).to(dev)
    M = torch.from_numpy(z['mid_c2x2'])
    D, U = eng.truncated_eig_sym(M.to(dev), chi)
    Dr, Ur = orc.truncated_eig_sym(M, chi)
    assert float((D.cpu() - Dr).abs().max()) < 1e-12
    nC, nT, Dv = eng.move_c4v(a, torch.from_numpy(z['mid_C']).to(dev), torch.from_numpy(z['mid_T']).to(dev), chi)
    assert H.maxrel(nC.cpu(), torch.from_numpy(z['mid_nC'])) < 1e-12
    assert H.maxrel(nT.abs().cpu(), torch.from_numpy(z['mid_nT']).abs()) < 1e-9
    Cc, Tc = torch.from_numpy(z['init_C']).to(dev), torch.from_numpy(z['init_T']).to(dev)
    for _ in range(meta['n_iter']):
        Cc, Tc, _ = eng.move_c4v(a, Cc, Tc, chi)
    assert H.maxrel(Cc.cpu(), torch.from_numpy(z['final_C'])) < 1e-10
    e = orc.energy_j1j2_c4v(a.cpu(), Cc.cpu(), Tc.cpu(), 1.0, meta['j2'])
    assert abs(e - float(z['energy'][0])) < 1e-10 * abs(float(z['energy'][0]))


def test_known_answer_rvb_c4v_on_gpu(eng, dev):
    """The reference's TestRVB known answer (-0.47684229 @1e-8) reproduced by the CUDA path."""
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V, init_env_c4v
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    from peps_torch_b200.config import CTMARGS
    z = np.load(H.GOLD + '/rvb_c4v_known_answer.npz')
    st = IPEPS_C4V(torch.from_numpy(z['site']).to(dev))
    env = ENV_C4V(int(z['chi'][0]), st)
    init_env_c4v(st, env)
    args = CTMARGS(); args.ctm_max_iter = 200
    ctmrg_c4v.run(st, env, ctm_args=args)
    e = orc.energy_j1j2_c4v(st.site().cpu(), env.get_C().cpu(), env.get_T().cpu(), 1.0, float(z['j2'][0]))
    assert abs(e - float(z['energy'][0])) < 1e-8


def test_known_answer_j1j2_2site_on_gpu(eng, dev):
    """TestCtmrg_States 2SITE known answer (-0.4434603770143078 @1e-6) through the CUDA path."""
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    from peps_torch_b200.ctm.generic import ctmrg
    z = np.load(H.GOLD + '/j1j2_2site_known_answer.npz')
    sites = H.golden_sites(z)
    st = IPEPS(H.to_dev(sites, dev), orc.v2s_2site, int(z['lX'][0]), int(z['lY'][0]))
    env = ENV(int(z['chi'][0]), st)
    init_env(st, env)
    e_prev = None
    for _ in range(30):
        for d in orc.DIRECTIONS:
            for _r in range(st.lX if d in (orc.LEFT, orc.RIGHT) else st.lY):
                ctmrg.ctm_MOVE(d, st, env)
        e = orc.energy_j1j2(sites, orc.v2s_2site, cpu(env.C), cpu(env.T), 1.0, float(z['j2'][0]))
        if e_prev is not None and abs(e - e_prev) < 1e-8:
            break
        e_prev = e
    assert abs(e - float(z['energy'][0])) < 1e-6


@pytest.mark.parametrize('family', ['A', 'B'])
def test_config2_size_against_live_oracle(eng, dev, family):
    """BASELINE config 2 (4SITE D=3 chi=48 float64): two full iterations, CUDA vs oracle."""
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    from peps_torch_b200.ctm.generic import ctmrg
    D, chi, iters = 3, 48, 2
    sites = orc.random_state_4site(D, family=family)
    C, T = orc.init_env(sites, orc.v2s_4site, chi)
    orc.run(sites, orc.v2s_4site, 2, 2, C, T, chi, iters)
    st = IPEPS(H.to_dev(sites, dev), orc.v2s_4site, 2, 2)
    env = ENV(chi, st)
    init_env(st, env)
    for _ in range(iters):
        for d in orc.DIRECTIONS:
            for _r in range(2):
                ctmrg.ctm_MOVE(d, st, env)
    assert H.spectra_diff(env.C, C) < 1e-9
    assert H.env_abs_diff(env.C, env.T, C, T) < 1.2e-8     # 3 x the reference-vs-reference floor of 4e-9
    e_gpu = orc.energy_j1j2(sites, orc.v2s_4site, cpu(env.C), cpu(env.T), 1.0, 0.3)
    e_cpu = orc.energy_j1j2(sites, orc.v2s_4site, C, T, 1.0, 0.3)
    assert abs(e_gpu - e_cpu) <= 1e-10 * abs(e_cpu)


@pytest.mark.parametrize('dt', [torch.float64, torch.complex128])
def test_rectangular_cell_with_direction_dependent_bonds(eng, dev, dt):
    """A 3 x 2 unit cell of six different sites whose vertical bonds (D = 2) differ from the horizontal ones (D = 3): the
    number of moves per direction follows lX / lY (ctmrg.py:88-96), the enlarged corners are rectangular in the bond
    dimensions, and up/down projectors have a different shape from left/right ones.  Two iterations against the live oracle."""
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    from peps_torch_b200.ctm.generic import ctmrg
    lX, lY, chi, Dv, Dh = 3, 2, 12, 2, 3
    g = torch.Generator().manual_seed(77)
    sites = OrderedDict()
    for y in range(lY):
        for x in range(lX):
            a = torch.randn(2, Dv, Dh, Dv, Dh, dtype=dt, generator=g)
            sites[(x, y)] = a / a.abs().max()

    def v2s(c):
        return (c[0] % lX, c[1] % lY)
    C, T = orc.init_env(sites, v2s, chi)
    orc.run(sites, v2s, lX, lY, C, T, chi, 2)
    st = IPEPS(H.to_dev(sites, dev), v2s, lX, lY)
    env = ENV(chi, st)
    init_env(st, env)
    for _ in range(2):
        for d in orc.DIRECTIONS:
            for _r in range(lX if d in (orc.LEFT, orc.RIGHT) else lY):
                ctmrg.ctm_MOVE(d, st, env)
    assert set(env.C) == set(C) and set(env.T) == set(T)
    for k in T:
        assert tuple(env.T[k].shape) == tuple(T[k].shape), k
    assert H.spectra_diff(env.C, C) < 1e-9
    assert H.env_abs_diff(env.C, env.T, C, T) < 1.2e-8


def _decaying(n, dt, dev, rate, seed):
    """n x n matrix with singular values rate^j (the CTM matrices M = R^T Rt decay geometrically)."""
    g = torch.Generator().manual_seed(seed)
    Q1, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, generator=g))
    Q2, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, generator=g))
    s = rate ** torch.arange(n, dtype=torch.float64)
    return ((Q1 * s.to(dt)) @ Q2.conj().t()).to(dev)


@pytest.mark.parametrize('shape', [(3, 48, torch.float64, 0.9), (4, 32, torch.complex128, 0.93), (8, 24, torch.float64, 0.97),
                                   (2, 64, torch.complex128, 0.9), (2, 61, torch.complex128, 0.9)])   # complex sketches 112 / 106 wide
def test_projector_biorthogonality_property(eng, dev, shape):
    """Size-independent property: Pt^T P = diag(1 on kept, 0 on cut) (ctm_projectors.py:279-293:
    P = R conj(U) S^-1/2, Pt = Rt V S^-1/2 with M = R^T Rt = U S V^H), at the full n of configs 2-4."""
    D, chi, dt, rate = shape
    n = chi * D * D
    R = _decaying(n, dt, dev, rate, 7)
    Rt = _decaying(n, dt, dev, rate, 8)
    P, Pt, S = eng.projectors(R, Rt, chi)
    Sr = torch.linalg.svdvals((R.t() @ Rt).cpu())[:chi]
    keep = int((Sr / Sr[0] > 1e-8).sum())
    G = (Pt.t() @ P).cpu()
    I = torch.zeros(chi, chi, dtype=dt); I[:keep, :keep] = torch.eye(keep, dtype=dt)
    assert float((G - I).abs().max()) < 1e-7
    assert float((S.cpu()[:keep] - Sr[:keep]).abs().max() / Sr[0]) < 1e-10


def test_flat_spectrum_keeps_biorthogonality(eng, dev):
    """A flat spectrum is the worst case for a range finder: the singular values are then only
    Ritz approximations, but the projector pair stays exactly bi-orthogonal by construction."""
    torch.manual_seed(7)
    n, chi = 432, 48
    R = torch.randn(n, n, dtype=torch.float64, device=dev) / n ** 0.5
    Rt = torch.randn(n, n, dtype=torch.float64, device=dev) / n ** 0.5
    P, Pt, S = eng.projectors(R, Rt, chi)
    assert H.maxrel((Pt.t() @ P).cpu(), torch.eye(chi, dtype=torch.float64)) < 1e-9


def test_truncated_svd_edge_cases(eng, dev):
    torch.manual_seed(11)
    # exact multiplet cut in half: the whole multiplet must be dropped (custom_svd.py:70-95)
    n, chi = 40, 6
    Q1, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64))
    Q2, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64))
    s = torch.cat([torch.tensor([1.0, 0.8, 0.6, 0.5, 0.3, 0.2, 0.2, 0.2]), 0.01 * torch.rand(n - 8)]).double()
    M = (Q1 * s) @ Q2.t()
    U, S, V = eng.truncated_svd(M.to(dev), chi)
    assert torch.allclose(S.cpu()[:5], s[:5], atol=1e-13) and float(S[5]) == 0.0
    assert float(U[:, 5].abs().max()) == 0.0 and float(V[:, 5].abs().max()) == 0.0
    # sign convention: the largest-|.| entry of every kept U column is positive (svd_gesdd.py:18-26)
    Uc = U.cpu()[:, :5]
    idx = Uc.abs().argmax(dim=0)
    assert bool((Uc[idx, torch.arange(5)] > 0).all())
    # rank-deficient input: trailing singular values come out (numerically) zero
    A = torch.randn(60, 3, dtype=torch.float64) @ torch.randn(3, 60, dtype=torch.float64)
    U, S, V = eng.truncated_svd(A.to(dev), 8)
    Sr = torch.linalg.svdvals(A)
    assert float((S.cpu()[:3] - Sr[:3]).abs().max() / Sr[0]) < 1e-13 and float(S[3:].abs().max() / Sr[0]) < 1e-12
    # chi == n: nothing to truncate
    B = torch.randn(12, 12, dtype=torch.float64)
    U, S, V = eng.truncated_svd(B.to(dev), 12)
    assert H.maxrel(((U * S) @ V.t()).cpu(), B) < 1e-12


def test_errors_are_python_exceptions(eng, dev):
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.config import CTMARGS
    z, meta = H.load_golden('generic_4site_D2_chi8_B')
    sites = H.golden_sites(z)
    C, T = H.golden_env(z, 'mid_')
    st = H.State(H.to_dev(sites, dev), orc.v2s_4site, 2, 2)
    env = H.Env(meta['chi'], H.to_dev(C, dev), H.to_dev(T, dev))
    with pytest.raises(ValueError):
        ctmrg.ctm_MOVE((1, 1), st, env)
    bad = CTMARGS(); bad.projector_method = '5X5'
    with pytest.raises(ValueError):
        ctmrg.ctm_MOVE((0, -1), st, env, ctm_args=bad)
    bad = CTMARGS(); bad.projector_svd_method = 'NOPE'
    with pytest.raises(TypeError):
        ctmrg.ctm_MOVE((0, -1), st, env, ctm_args=bad)
    with pytest.raises(TypeError):
        eng.einsum2('ab,bc->ac', torch.zeros(2, 2, device=dev, dtype=torch.float16),
                    torch.zeros(2, 2, device=dev, dtype=torch.float16))       # only float64/complex128 (+ widened float32/complex64)
    # float32 operands are widened on entry and rounded on exit (north star: 1e-4 gate)
    a32 = torch.rand(5, 7, device=dev); b32 = torch.rand(7, 3, device=dev)
    c32 = eng.einsum2('ab,bc->ac', a32, b32)
    assert c32.dtype == torch.float32 and float((c32 - a32 @ b32).abs().max()) < 1e-5


def _graded(n, dt, decades, seed):
    """n x n matrix with geometrically decaying singular values sigma_i = 10^(-decades*i/n)."""
    g = torch.Generator().manual_seed(seed)
    U, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, generator=g))
    V, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, generator=g))
    s = torch.logspace(0, -decades, n, dtype=torch.float64)
    return (U * s.to(dt)) @ V.conj().t(), s


@pytest.mark.parametrize('n,chi,dt', [(1300, 100, torch.float64), (800, 70, torch.complex128)])
def test_large_sketch_blocked_qr_and_multi_cta_jacobi(eng, dev, n, chi, dt):
    """Sizes beyond the register / single-cluster QR and the shared-memory Jacobi (the decomposition path of
    configs c3 and c5): sketch k = 2 chi > 128 columns -> blocked Householder QR, k x k Jacobi over several CTAs.
    Checked against the full LAPACK SVD: singular values to 1e-10 relative, U/V through the rank-chi projector."""
    M, s = _graded(n, dt, 12.0, 5)
    U, S, V = eng.truncated_svd(M.to(dev), chi)      # residual-checked range finder (library default)
    Ur, Sr, Vhr = torch.linalg.svd(M)
    S, U, V = S.cpu(), U.cpu(), V.cpu()
    assert float(((S - Sr[:chi]).abs() / Sr[:chi]).max()) < 1e-10
    eye = torch.eye(chi, dtype=dt)
    assert float((U.conj().t() @ U - eye).abs().max()) < 1e-12
    assert float((V.conj().t() @ V - eye).abs().max()) < 1e-12
    best = (Ur[:, :chi] * Sr[:chi].to(dt)) @ Vhr[:chi]
    rec = (U * S.to(dt)) @ V.conj().t()
    assert H.maxrel(rec, best) < 1e-10


def test_large_hermitian_eig_multi_cta(eng, dev):
    n, chi = 900, 80
    g = torch.Generator().manual_seed(11)
    Q, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.complex128, generator=g))
    lam = torch.logspace(0, -5, n, dtype=torch.float64) * torch.where(torch.arange(n) % 3 == 0, -1.0, 1.0)
    M = (Q * lam.to(Q.dtype)) @ Q.conj().t()
    M = 0.5 * (M + M.conj().t())
    D, U = eng.truncated_eig_sym(M.to(dev), chi)
    Dr, Ur = orc.truncated_eig_sym(M, chi)
    assert float(((D.cpu() - Dr).abs() / Dr.abs()).max()) < 1e-10
    U = U.cpu()
    assert float((U.conj().t() @ U - torch.eye(chi, dtype=U.dtype)).abs().max()) < 1e-12
    assert H.maxrel((U * D.cpu().to(U.dtype)) @ U.conj().t(), (Ur * Dr.to(Ur.dtype)) @ Ur.conj().t()) < 1e-9


@pytest.mark.parametrize('kind', ['LU', 'RU', 'RD', 'LD'])
def test_fused_double_layer_corner_D8(eng, dev, kind):
    """D = 8, p = 2 real corners take the fused double-layer kernel (dl_fused.cu): checked element-wise
    against the oracle's restatement of c2x2_*_sl_c on the same random inputs (deterministic: 1e-13)."""
    D, chi, p = 8, 12, 2
    g = torch.Generator().manual_seed(7)
    a = torch.randn(p, D, D, D, D, dtype=torch.float64, generator=g)
    C = torch.randn(chi, chi, dtype=torch.float64, generator=g)
    shapes = {(0, -1): (chi, D * D, chi), (-1, 0): (chi, chi, D * D), (0, 1): (D * D, chi, chi), (1, 0): (chi, D * D, chi)}
    kc, k1, k2, _ = orc.CORNERS[kind]
    T1 = torch.randn(shapes[k1], dtype=torch.float64, generator=g)
    T2 = torch.randn(shapes[k2], dtype=torch.float64, generator=g)
    ref = orc.c2x2(kind, C, T1, T2, a)
    out = eng.c2x2(kind, C.to(dev), T1.to(dev), T2.to(dev), a.to(dev), chi)
    assert out.shape == ref.shape
    assert H.maxrel(out.cpu(), ref) < 1e-13


def test_tall_sketch_two_level_blocked_qr(eng, dev):
    """4096 rows: a cluster holds only 32-column leaf panels, so the sketch (k = 200) goes through the two-level
    blocked QR (leaf panels inside 128-column super panels) and the multi-CTA Jacobi.  The matrix is built from
    known factors, so no reference SVD of a 4096^2 matrix is needed: sigma to 1e-10, singular vectors by overlap."""
    n, chi = 4096, 100
    g = torch.Generator(device='cpu').manual_seed(3)
    U0, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, generator=g).to(dev))
    V0, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, generator=g).to(dev))
    s = torch.logspace(0, -40, n, dtype=torch.float64, device=dev)
    M = (U0 * s) @ V0.t()
    U, S, V = eng.truncated_svd(M, chi)
    assert float(((S - s[:chi]).abs() / s[:chi]).max()) < 1e-10
    eye = torch.eye(chi, dtype=torch.float64, device=dev)
    assert float((U.t() @ U - eye).abs().max()) < 1e-12
    assert float((V.t() @ V - eye).abs().max()) < 1e-12
    keep = s[:chi] / s[0] > 1e-8
    ou = (U0[:, :chi].t() @ U).diagonal().abs()[keep]
    ov = (V0[:, :chi].t() @ V).diagonal().abs()[keep]
    assert float((1 - ou).max()) < 1e-9 and float((1 - ov).max()) < 1e-9


@pytest.mark.parametrize('name', ['generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'])
def test_matrix_free_projectors_match_reference_fixtures(eng, dev, name):
    """The matrix-free projector path (M = R^T Rt applied as four corner factors, used when D^2 > 20: config c5)
    forced on the small fixtures: every move must reproduce the reference's outputs like the explicit path."""
    from peps_torch_b200 import _lib
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'mid_')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    eng.debug_set_matrix_free(2)
    try:
        for d in orc.DIRECTIONS:
            env = H.Env(chi, H.to_dev(C, dev), H.to_dev(T, dev))
            eng.move_generic(d, st, env)
            Cg, Tg = H.golden_env(z, f'move_{d[0]}_{d[1]}_')
            assert H.env_abs_diff(env.C, env.T, Cg, Tg) < 1e-8
    finally:
        eng.debug_set_matrix_free(1)


def test_slowly_decaying_spectrum_against_live_oracle(eng, dev):
    """4SITE D=4 chi=40 (family B): chi cuts through a slowly decaying spectrum of M.  A fixed four power
    iterations left range-finder residuals of 2e-12 here and |C|,|T| 3e-7 / spectra 2e-9 away from the oracle (the
    reference's own gesdd-vs-gesvd floor is 1.3e-9 / 1.4e-12); the residual-checked default must be at that floor."""
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    from peps_torch_b200.ctm.generic import ctmrg
    D, chi, iters = 4, 40, 3
    sites = orc.random_state_4site(D, family='B')
    C, T = orc.init_env(sites, orc.v2s_4site, chi)
    orc.run(sites, orc.v2s_4site, 2, 2, C, T, chi, iters)
    st = IPEPS(H.to_dev(sites, dev), orc.v2s_4site, 2, 2)
    env = ENV(chi, st)
    init_env(st, env)
    for _ in range(iters):
        for d in orc.DIRECTIONS:
            for _r in range(2):
                ctmrg.ctm_MOVE(d, st, env)
    assert H.spectra_diff(env.C, C) < 1e-10
    assert H.env_abs_diff(env.C, env.T, C, T) < 1.2e-8
    e_gpu = orc.energy_j1j2(sites, orc.v2s_4site, cpu(env.C), cpu(env.T), 1.0, 0.3)
    e_cpu = orc.energy_j1j2(sites, orc.v2s_4site, C, T, 1.0, 0.3)
    assert abs(e_gpu - e_cpu) <= 1e-12 * abs(e_cpu)


@pytest.mark.parametrize('name', ['generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128'])
def test_projector_method_4x2_against_oracle(eng, dev, name):
    """CTMARGS.projector_method='4X2' (ctm_projectors.py:66-136): R, Rt are single enlarged corners.  The oracle's
    4X2 restatement is pinned against the reference in tests/test_oracle_vs_reference_cpu.py."""
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'mid_')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    for d in orc.DIRECTIONS:
        C, T = dict(C0), dict(T0)
        orc.ctm_move(d, sites, v2s, C, T, chi, orc.OracleArgs(projector_method='4X2'))
        env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
        eng.move_generic(d, st, env, projector_method=1)
        assert H.env_abs_diff(env.C, env.T, C, T) < 1e-8


# ------------------------------------------------------------------------------------------
# non-default variants of the move: the oracle restatements used below are pinned against the
# reference in tests/test_oracle_vs_reference_cpu.py
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'])
def test_double_layer_sites_against_oracle(eng, dev, name):
    """Rank-4 on-site tensors A = a (x) a* (ctm_force_dl, ctmrg.py:51-61): corners element-wise, moves through |.|."""
    from peps_torch_b200.ctm.generic import ctmrg
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'mid_')
    dl_gpu = type(sites)((c, ctmrg.double_layer(eng, a.to(dev))) for c, a in sites.items())
    dl = type(sites)((c, orc.double_layer(a)) for c, a in sites.items())
    for c in sites:
        assert H.maxrel(dl_gpu[c].cpu(), dl[c]) < 1e-14
    coord = list(sites.keys())[-1]
    for kind in orc.CORNERS:
        kc, k1, k2, _ = orc.CORNERS[kind]
        out = eng.c2x2(kind, C0[(coord, kc)].to(dev), T0[(coord, k1)].to(dev), T0[(coord, k2)].to(dev), dl_gpu[coord], chi)
        assert H.maxrel(out.cpu(), torch.from_numpy(z['c2x2_' + kind])) < 1e-13
    st = H.State(dl_gpu, v2s, lX, lY)
    for d in orc.DIRECTIONS:
        C, T = dict(C0), dict(T0)
        orc.ctm_move(d, dl, v2s, C, T, chi)
        env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
        eng.move_generic(d, st, env)
        assert H.env_abs_diff(env.C, env.T, C, T) < 1e-8


@pytest.mark.parametrize('name', ['generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128'])
def test_2norm_normalisation_against_oracle(eng, dev, name):
    """ctm_absorb_normalization != 'inf' -> vector 2-norm (ctmrg.py:210-230)."""
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.config import CTMARGS
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'mid_')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    args = CTMARGS(); args.ctm_absorb_normalization = 'fro'
    for d in orc.DIRECTIONS:
        C, T = dict(C0), dict(T0)
        orc.ctm_move(d, sites, v2s, C, T, chi, orc.OracleArgs(ctm_absorb_normalization='fro'))
        env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
        ctmrg.ctm_MOVE(d, st, env, ctm_args=args)
        assert H.env_abs_diff(env.C, env.T, C, T) < 1e-8
        kC1, kC2, kT = orc.ABSORB[d]['out']
        for k, t in list(env.C.items()) + list(env.T.items()):
            if k[1] in (kC1, kC2, kT):
                assert abs(float(torch.linalg.vector_norm(t)) - 1.0) < 1e-13


def test_run_overlap_against_oracle(eng, dev):
    """run_overlap (ctmrg.py:112-175): double-layer tensors ket = state1, bra = state2 (not Hermitian)."""
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS
    z, meta = H.load_golden('generic_4site_D2_chi8_B')
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'mid_')
    g = torch.Generator().manual_seed(7)
    sites2 = type(sites)((c, a + 0.05 * (torch.rand(a.shape, generator=g, dtype=a.dtype) - 0.5)) for c, a in sites.items())
    args = CTMARGS(); args.ctm_force_dl = True; args.ctm_max_iter = 2
    env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
    s1, s2 = IPEPS(H.to_dev(sites, dev), v2s, lX, lY), IPEPS(H.to_dev(sites2, dev), v2s, lX, lY)
    calls = []
    ctmrg.run_overlap(s1, s2, env, conv_check=lambda a, b, e, h, ctm_args=None: (calls.append(1) or False, h), ctm_args=args)
    assert len(calls) == 2
    with pytest.raises(AssertionError):
        ctmrg.run_overlap(s1, s2, env, ctm_args=CTMARGS())
    dl = type(sites)((c, orc.double_layer(sites[c], sites2[c])) for c in sites)
    C, T = dict(C0), dict(T0)
    for _ in range(2):
        for d in orc.DIRECTIONS:
            orc.ctm_move(d, dl, v2s, C, T, chi)
    assert H.env_abs_diff(env.C, env.T, C, T) < 1e-8


def test_run_force_dl_and_warmup_match_plain_run(eng, dev):
    """ctmrg.run with ctm_force_dl (ctmrg.py:51-61) and with warm-up iterations (:76-86) against the oracle."""
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS
    z, meta = H.load_golden('generic_4site_D2_chi8_B')
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C0, T0 = H.golden_env(z, 'init_')
    st = IPEPS(H.to_dev(sites, dev), v2s, lX, lY)
    C, T = dict(C0), dict(T0)
    for _ in range(3):
        orc.ctm_iteration(sites, v2s, lX, lY, C, T, chi)
    e_ref = orc.energy_j1j2(sites, v2s, C, T, 1.0, 0.3)
    for kw in (dict(ctm_force_dl=True, ctm_max_iter=3), dict(ctm_warmup_iter=1, ctm_max_iter=1)):
        args = CTMARGS()
        for k, v in kw.items():
            setattr(args, k, v)
        env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
        _, _, t_ctm, _ = ctmrg.run(st, env, ctm_args=args)
        assert t_ctm > 0
        # warm-up: max(ctm_warmup_iter, ceil(chi / D^2)) = 2 iterations, then ctm_max_iter = 1
        assert H.env_abs_diff(env.C, env.T, C, T) < 1e-8
        e = orc.energy_j1j2(sites, v2s, cpu(env.C), cpu(env.T), 1.0, 0.3)
        assert abs(e - e_ref) < 1e-10 * abs(e_ref)


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_c4v_double_layer_move_and_2norm(eng, dev, name):
    """ctm_MOVE_dl / run_dl (ctmrg_c4v.py:110-176,200-322) and the 2-norm branch of _move_normalize_c (:182-197)."""
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    z, meta = H.load_golden(name)
    chi = meta['chi']
    a = torch.from_numpy(z['site'])
    A = orc.double_layer(a)
    C0, T0 = torch.from_numpy(z['init_C']), torch.from_numpy(z['init_T'])
    stc = IPEPS_C4V(a.to(dev))
    for norm in ('inf', 'fro'):
        args = CTMARGS(); args.ctm_absorb_normalization = norm; args.ctm_max_iter = 3
        C, T = C0, T0
        for _ in range(3):
            C, T = orc.ctm_move_c4v(A, C, T, chi, orc.OracleArgs(ctm_absorb_normalization=norm))
        # run_dl from the single-layer tensor (A built once on the GPU)
        env = ENV_C4V(chi, stc)
        env.C[env.keyC], env.T[env.keyT] = C0.to(dev), T0.to(dev)
        ctmrg_c4v.run_dl(stc, env, ctm_args=args)
        assert H.maxrel(env.get_C().cpu(), C) < 1e-10
        assert H.maxrel(env.get_T().abs().cpu(), T.abs()) < 1e-8
        # ctm_MOVE_dl handed the double-layer tensor itself
        env2 = ENV_C4V(chi, stc)
        env2.C[env2.keyC], env2.T[env2.keyT] = C0.to(dev), T0.to(dev)
        for _ in range(3):
            ctmrg_c4v.ctm_MOVE_dl(A.to(dev), env2, None, ctm_args=args)
        assert H.maxrel(env2.get_C().cpu(), C) < 1e-10
        assert H.maxrel(env2.get_T().abs().cpu(), T.abs()) < 1e-8
        if norm == 'fro':
            assert abs(float(torch.linalg.vector_norm(env2.get_T())) - 1.0) < 1e-13
    with pytest.raises(ValueError):
        ctmrg_c4v.ctm_MOVE_sl(A.to(dev), env2)


# ------------------------------------------------------------------------------------------
# reduced density matrices / energy on the converged environment
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,dt', [(4, torch.float64), (4, torch.complex128), (16, torch.float64), (16, torch.complex128),
                                  (64, torch.float64), (64, torch.complex128)])
def test_sym_pos_def_matrix_against_oracle(eng, dev, n, dt):
    """_sym_pos_def_matrix (ctm/generic/rdm.py:38-57) on an indefinite, slightly non-Hermitian matrix."""
    g = torch.Generator().manual_seed(n)
    Q, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, generator=g))
    d = torch.linspace(1.0, -0.2, n, dtype=torch.float64)
    M = (Q * d.to(dt)) @ Q.conj().t() + 1e-3 * torch.randn(n, n, dtype=dt, generator=g)
    for spd in (False, True):
        out = eng.sym_pos_def(M.to(dev), spd)
        assert H.maxrel(out.cpu(), orc._sym_pos_def(M, spd)) < 1e-12
    # nothing negative: sym_pos_def=True leaves the hermitised matrix alone (rdm.py:49)
    P = (Q * torch.linspace(1.0, 0.01, n, dtype=torch.float64).to(dt)) @ Q.conj().t()
    assert H.maxrel(eng.sym_pos_def(P.to(dev), True).cpu(), orc._sym_pos_def(P, True)) < 1e-13
    # a 4-site RDM-shaped tensor keeps its shape
    if n == 16:
        t = eng.sym_pos_def(M.reshape((2,) * 8).to(dev), True)
        assert t.shape == (2,) * 8 and H.maxrel(t.cpu().reshape(16, 16), orc._sym_pos_def(M, True)) < 1e-12


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128'])
def test_rdm2x2_and_energy_against_oracle(eng, dev, name):
    from peps_torch_b200.ctm.generic import rdm
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'mid_')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    env = H.Env(chi, H.to_dev(C, dev), H.to_dev(T, dev))
    hp = orc.j1j2_hp(1.0, 0.3, dtype=sites[(0, 0)].dtype)
    e_gpu = 0.
    for coord in sites:
        r_ref = orc.rdm2x2(coord, sites, v2s, C, T)
        r = rdm.rdm2x2(coord, st, env)
        assert r.shape == r_ref.shape and float((r.cpu() - r_ref).abs().max()) < 1e-12
        e_gpu = e_gpu + torch.einsum('ijklabcd,ijklabcd', r.cpu(), hp)
    e_gpu = float((e_gpu / len(sites)).real)
    e_ref = orc.energy_j1j2(sites, v2s, C, T, 1.0, 0.3)
    assert abs(e_gpu - e_ref) < 1e-12 * max(1.0, abs(e_ref))
    coord = list(sites)[1]
    for open_sites in ([0, 1], [0, 3], [2], [1, 2, 3]):
        for spd in (False, True):
            r_ref = orc.rdm2x2(coord, sites, v2s, C, T, open_sites=tuple(open_sites), sym_pos_def=spd)
            r = rdm.rdm2x2(coord, st, env, open_sites=open_sites, sym_pos_def=spd)
            assert r.shape == r_ref.shape and float((r.cpu() - r_ref).abs().max()) < 1e-12, (open_sites, spd)
    with pytest.raises(ValueError):
        rdm.rdm2x2(coord, st, env, open_sites=[])


@pytest.mark.parametrize('name', C4V)
def test_c4v_rdms_and_energy_against_reference_fixture(eng, dev, name):
    from peps_torch_b200.ctm.one_site_c4v import rdm_c4v
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    Cc, Tc = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    stc = IPEPS_C4V(a.to(dev))
    envc = ENV_C4V(meta['chi'], stc)
    envc.C[envc.keyC], envc.T[envc.keyT] = Cc.to(dev), Tc.to(dev)
    for spd in (False, True):
        assert float((rdm_c4v.rdm2x2_NN_lowmem_sl(stc, envc, sym_pos_def=spd).cpu() - orc.rdm2x2_c4v(a, Cc, Tc, (0, 1), spd)).abs().max()) < 1e-12
        assert float((rdm_c4v.rdm2x2_NNN_lowmem_sl(stc, envc, sym_pos_def=spd).cpu() - orc.rdm2x2_c4v(a, Cc, Tc, (0, 3), spd)).abs().max()) < 1e-12
    assert float((rdm_c4v.rdm2x2(stc, envc).cpu() - orc.rdm2x2_c4v(a, Cc, Tc)).abs().max()) < 1e-12
    # energy_1x1_lowmem (models/j1j2.py:641-679) from the GPU density matrices against the reference's own number
    sz, sp, sm, I = orc.spin_half_ops(a.dtype)
    SS = torch.einsum('ij,ab->iajb', sz, sz) + 0.5 * (torch.einsum('ij,ab->iajb', sp, sm) + torch.einsum('ij,ab->iajb', sm, sp))
    rot = torch.tensor([[0., 1.], [-1., 0.]], dtype=a.dtype)
    SS_rot = torch.einsum('ki,kjcb,ca->ijab', rot, SS, rot)
    e = 2.0 * torch.einsum('ijab,ijab', rdm_c4v.rdm2x2_NN_lowmem_sl(stc, envc, sym_pos_def=True).cpu(), SS_rot)
    if abs(meta['j2']) > 0:
        e = e + 2.0 * meta['j2'] * torch.einsum('ijab,ijab', rdm_c4v.rdm2x2_NNN_lowmem_sl(stc, envc, sym_pos_def=True).cpu(), SS)
    e_ref = float(z['energy'][0])
    assert abs(float(e.real) - e_ref) < 1e-10 * abs(e_ref)


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D3_chi12_B', 'generic_4site_D2_chi8_B_c128',
                                  'kagome_1site_D2_chi8_A'])
def test_small_rdms_against_oracle(eng, dev, name):
    """rdm1x1 / rdm2x1 / rdm1x2 (ctm/generic/rdm.py:71-112,304-350,622-670): deterministic contractions, element-wise."""
    from peps_torch_b200.ctm.generic import rdm
    z, meta = H.load_golden(name)
    chi = meta['chi']
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'mid_')
    st = H.State(H.to_dev(sites, dev), v2s, lX, lY)
    env = H.Env(chi, H.to_dev(C, dev), H.to_dev(T, dev))
    for coord in sites:
        for f, g in ((rdm.rdm1x1, orc.rdm1x1), (rdm.rdm2x1, orc.rdm2x1), (rdm.rdm1x2, orc.rdm1x2)):
            for spd in (False, True):
                if spd and sites[coord].shape[0] ** 2 > 160 and f is not rdm.rdm1x1:
                    continue                                   # kagome p = 8: 64 x 64 is fine, 4096 x 4096 is not needed
                r_ref = g(coord, sites, v2s, C, T, sym_pos_def=spd)
                r = f(coord, st, env, sym_pos_def=spd)
                # sym_pos_def=True: the positive projection goes through libctmb's Jacobi eigensolver (angle-based sweep
                # exit since round 2: no fixture-specific bound)
                tol = 1e-10 if spd else 1e-12
                assert r.shape == r_ref.shape and float((r.cpu() - r_ref).abs().max()) < tol, (f.__name__, coord, spd)
        # operator= : the unnormalised expectation value over the raw one-site network (rdm.py:89-90,175-181)
        p = sites[coord].shape[0]
        op = torch.randn(p, p, dtype=sites[coord].dtype, generator=torch.Generator().manual_seed(3))
        want = (orc.rdm1x1(coord, sites, v2s, C, T, raw=True) * op.t()).sum()
        got = rdm.rdm1x1(coord, st, env, operator=op.to(dev))
        assert abs(complex(got.cpu()) - complex(want)) < 1e-12 * abs(complex(want)), (coord, got, want)


@pytest.mark.parametrize('name', C4V)
def test_c4v_small_rdms_against_oracle(eng, dev, name):
    from peps_torch_b200.ctm.one_site_c4v import rdm_c4v
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    Cc, Tc = torch.from_numpy(z['final_C']), torch.from_numpy(z['final_T'])
    stc = IPEPS_C4V(a.to(dev))
    envc = ENV_C4V(meta['chi'], stc)
    envc.C[envc.keyC], envc.T[envc.keyT] = Cc.to(dev), Tc.to(dev)
    for spd in (False, True):
        assert float((rdm_c4v.rdm1x1_sl(stc, envc, sym_pos_def=spd).cpu() - orc.rdm_small_c4v('1x1', a, Cc, Tc, spd)).abs().max()) < 1e-12
        assert float((rdm_c4v.rdm2x1_sl(stc, envc, sym_pos_def=spd).cpu() - orc.rdm_small_c4v('2x1', a, Cc, Tc, spd)).abs().max()) < 1e-12
        assert float((rdm_c4v.rdm3x1_sl(stc, envc, sym_pos_def=spd).cpu() - orc.rdm3x1_c4v(a, Cc, Tc, spd)).abs().max()) < 1e-12


# ------------------------------------------------------------------------------------------
# libctmb against outputs of the UNMODIFIED reference for the variants and the density matrices
# (tests/golden/variants_*.npz, oracle/gen_golden_variants.py)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128'])
def test_variants_against_reference_fixtures_generic(eng, dev, name):
    # (bounds: libctmb-vs-oracle of the tests above plus oracle-vs-reference of the generator, 1.4e-10 / 1e-13)
    assert H.check_generic_variants(name, dev, tol_move=1e-8, tol_rdm=5e-12, tol_rdm_spd=1e-10) == 3 * 4 * 12 + 4 * 2 * 4


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_variants_against_reference_fixtures_c4v(eng, dev, name):
    assert H.check_c4v_variants(name, dev, tol_C=1e-10, tol_T=1e-8, tol_rdm=5e-12, tol_rdm_spd=1e-10) == 12


def test_config1_script_known_answer_on_gpu(eng, dev):
    """BASELINE.json configs[0] (J1-J2 one-site C4v D=2 chi=16 float64): the reference script
    `ctmrg_j1j2_c4v.py --bond_dim 2 --chi 16 --seed 123 --j2 0.3` stops after four moves and prints
    FINAL -0.35003258049356745.  Same state, four moves through the drop-in `run`, energy_1x1_lowmem
    (models/j1j2.py:641-679) from the GPU density matrices."""
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v, rdm_c4v
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V, init_env_c4v
    a = orc.random_state_c4v(2, family='A')
    stc = IPEPS_C4V(a.to(dev))
    env = ENV_C4V(16, stc)
    init_env_c4v(stc, env)
    args = CTMARGS(); args.ctm_max_iter = 4
    ctmrg_c4v.run(stc, env, ctm_args=args)
    sz, sp, sm, I = orc.spin_half_ops(a.dtype)
    SS = torch.einsum('ij,ab->iajb', sz, sz) + 0.5 * (torch.einsum('ij,ab->iajb', sp, sm) + torch.einsum('ij,ab->iajb', sm, sp))
    rot = torch.tensor([[0., 1.], [-1., 0.]], dtype=a.dtype)
    SS_rot = torch.einsum('ki,kjcb,ca->ijab', rot, SS, rot)
    e = 2.0 * torch.einsum('ijab,ijab', rdm_c4v.rdm2x2_NN_lowmem_sl(stc, env, sym_pos_def=True).cpu(), SS_rot) \
        + 2.0 * 0.3 * torch.einsum('ijab,ijab', rdm_c4v.rdm2x2_NNN_lowmem_sl(stc, env, sym_pos_def=True).cpu(), SS)
    assert abs(float(e) - (-0.35003258049356745)) < 1e-10 * 0.35


def test_config2_script_known_answer_on_gpu(eng, dev):
    """BASELINE.json configs[1] (J1-J2 generic 4SITE D=3 chi=48 float64) on the script-exact state: the reference script
    `ctmrg_j1j2.py --tiling 4SITE --bond_dim 3 --chi 48 --seed 123 --j2 0.3` converges in three iterations and prints
    FINAL 0.6424192641900255 (energy per site, models/j1j2.py:223-247).  Same state, three iterations through the drop-in
    `run`, energy from the GPU rdm2x2 of every plaquette."""
    from peps_torch_b200.ctm.generic import ctmrg, rdm
    from peps_torch_b200.config import CTMARGS
    from peps_torch_b200.ipeps import IPEPS
    from peps_torch_b200.env import ENV, init_env
    sites = orc.random_state_4site(3, family='A')
    st = IPEPS(H.to_dev(sites, dev), orc.v2s_4site, 2, 2)
    env = ENV(48, st)
    init_env(st, env)
    args = CTMARGS(); args.ctm_max_iter = 3
    ctmrg.run(st, env, ctm_args=args)
    hp = orc.j1j2_hp(1.0, 0.3)
    e = sum(torch.einsum('ijklabcd,ijklabcd', rdm.rdm2x2(c, st, env).cpu(), hp) for c in sites) / len(sites)
    assert abs(float(e) - 0.6424192641900255) < 1e-10 * 0.64


@pytest.mark.parametrize('lag', [0, 1])
def test_conv_rdm2x1_criterion_on_gpu(eng, dev, lag):
    """ctmrg_conv_rdm2x1 (examples/j1j2/ctmrg_j1j2_c4v.py:101-129) with libctmb moves and density matrices on the config-1
    state: the distances the unmodified script prints on CPU; lag=1 is the variant without a stream stall."""
    import test_conv_rdm_cpu as T
    T.check(T.run_config1(dev, lag), lag)


@pytest.mark.parametrize('dt', [torch.float64, torch.complex128])
def test_householder_qr_entry_point(eng, dev, dt):
    """ctmb_qr against torch.linalg.qr (same LAPACK sign convention: element-wise), through all drivers: register kernel
    (432 x 84), WY form (1024 x 128), blocked factorisation (1536 x 96, the C.T matrix of config 3; 4096 x 200); square
    matrices; widths around the shared-memory bound of the WY solve (complex: k = 105 is the last WY width, 106-113 used to
    pass the support test and then fail at launch)."""
    for rows, k in ((64, 16), (432, 84), (1024, 128), (1536, 96), (4096, 200), (108, 108), (420, 105), (420, 106), (339, 113),
                    (128, 128)):
        g = torch.Generator().manual_seed(rows + k)
        M = torch.randn(rows, k, dtype=dt, generator=g)
        Q, R = eng.qr(M.to(dev))
        Q, R = Q.cpu(), R.cpu()
        Qr, Rr = torch.linalg.qr(M)
        eye = torch.eye(k, dtype=dt)
        assert float((Q.conj().t() @ Q - eye).abs().max()) < 1e-13, (rows, k)
        assert H.maxrel(Q @ R, M) < 1e-13, (rows, k)
        assert float(R.tril(-1).abs().max()) == 0.0
        # unique up to the phase of each column of Q / row of R.  LAPACK makes the diagonal of R real; so does libctmb, except
        # for the last column of a SQUARE complex matrix (no entries below the diagonal: no reflector is generated and
        # R[k-1,k-1] keeps its phase).  Compare after aligning the phases: Q_ours diag(ph) = Q_lapack.
        ph = (R.diagonal() / Rr.diagonal())
        ph = ph / ph.abs()
        if dt.is_complex and rows != k:
            assert float(ph.imag.abs().max()) < 1e-12, (rows, k)          # real diagonal, as LAPACK's
        assert H.maxrel(Q * ph[None, :], Qr) < 1e-11, (rows, k)


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_c4v_qr_move_against_oracle(eng, dev, name):
    """ctm_MOVE_QR_sl (ctmrg_c4v.py:465-602) composed from libctmb calls against the oracle (pinned element-wise against the
    reference by tests/test_oracle_vs_reference_cpu.py); also at the size of config 3 (D = 4, chi = 96, complex)."""
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    from peps_torch_b200.env import ENV_C4V
    from peps_torch_b200.ipeps import IPEPS_C4V
    z, meta = H.load_golden(name)
    a = torch.from_numpy(z['site'])
    cases = [(a, meta['chi'])]
    if a.is_complex():
        cases.append((orc.random_state_c4v(4, family='B', dtype=torch.complex128), 96))
    for a, chi in cases:
        C, T = orc.init_env_c4v(a, chi)
        for _ in range(2):
            C, T = orc.ctm_move_c4v(a, C, T, chi)
        env = ENV_C4V(chi, IPEPS_C4V(a.to(dev)))
        env.C[env.keyC], env.T[env.keyT] = C.to(dev), T.to(dev)
        for _ in range(2):
            C, T = orc.ctm_move_qr_c4v(a, C, T, chi)
            ctmrg_c4v.ctm_MOVE_QR_sl(a.to(dev), env)
            assert H.maxrel(env.C[env.keyC].abs().cpu(), C.abs()) < 1e-11
            assert H.maxrel(env.T[env.keyT].abs().cpu(), T.abs()) < 1e-11


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'])
def test_transfer_operator_matvecs_and_spectra(eng, dev, name, monkeypatch):
    """apply_TM_1sO / apply_TM_0sO and get_Top_spec / get_Top_w0_spec (ctm/generic/corrf.py, transferops.py) on libctmb against
    the same functions with the oracle as engine (which tests/test_transferops_cpu.py pins against the unmodified reference)."""
    from peps_torch_b200.ctm.generic import transferops as ot, corrf as oc
    from test_transferops_cpu import edge_shapes, DIRS
    z, meta = H.load_golden(name)
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'final_' if any(k.startswith('final_') for k in z.files) else 'mid_')
    dt = next(iter(sites.values())).dtype
    st_c, env_c = H.State(sites, v2s, lX, lY), H.Env(meta['chi'], dict(C), dict(T))
    st_g, env_g = H.State(H.to_dev(sites, dev), v2s, lX, lY), H.Env(meta['chi'], H.to_dev(C, dev), H.to_dev(T, dev))
    oracle = H.OracleEngine()
    for d in DIRS:
        g = torch.Generator().manual_seed(1)
        chi1, d2, chi2 = edge_shapes(st_c, env_c, d)
        V = torch.randn(chi1, d2, chi2, dtype=dt, generator=g)
        V0 = torch.randn(chi1, chi2, dtype=dt, generator=g)
        monkeypatch.setattr(oc, '_engine', lambda: oracle)
        want1, want0 = oc.apply_TM_1sO((0, 0), d, st_c, env_c, V), oc.apply_TM_0sO((0, 0), d, st_c, env_c, V0)
        Lw, Ww = ot.get_Top_spec(3, (0, 0), d, st_c, env_c), ot.get_Top_w0_spec(3, (0, 0), d, st_c, env_c)
        monkeypatch.setattr(oc, '_engine', lambda: eng)
        assert H.maxrel(oc.apply_TM_1sO((0, 0), d, st_g, env_g, V.to(dev)).cpu(), want1) < 1e-13
        assert H.maxrel(oc.apply_TM_0sO((0, 0), d, st_g, env_g, V0.to(dev)).cpu(), want0) < 1e-13
        assert float((ot.get_Top_spec(3, (0, 0), d, st_g, env_g).cpu() - Lw).abs().max()) < 1e-10
        assert float((ot.get_Top_w0_spec(3, (0, 0), d, st_g, env_g).cpu().abs() - Ww.abs()).abs().max()) < 1e-10
        if lX == lY == 1:           # entanglement spectrum of the L-leg cylinder (transferops.py:207-370; one-site cells)
            from test_transferops_cpu import _spec
            monkeypatch.setattr(oc, '_engine', lambda: oracle)
            Sw = ot.get_EH_spec_Ttensor(2, 3, (0, 0), d, st_c, env_c)
            monkeypatch.setattr(oc, '_engine', lambda: eng)
            assert float((_spec(ot.get_EH_spec_Ttensor(2, 3, (0, 0), d, st_g, env_g).cpu()) - _spec(Sw)).abs().max()) < 1e-9


@pytest.mark.parametrize('name', ['generic_4site_D2_chi8_B', 'generic_4site_D2_chi8_B_c128', 'kagome_1site_D2_chi8_A'])
def test_edges_operator_insertions_and_corrf(eng, dev, name, monkeypatch):
    """get_edge / apply_edge / apply_TM_1sO with a one-site operator / corrf_1sO1sO (ctm/generic/corrf.py:10-104, 234-277,
    415-419, 980-1067) on libctmb against the same functions with the oracle as engine (pinned against the unmodified
    reference by tests/test_transferops_cpu.py)."""
    from peps_torch_b200.ctm.generic import corrf as oc
    from test_transferops_cpu import edge_shapes, DIRS
    z, meta = H.load_golden(name)
    sites = H.golden_sites(z)
    v2s, lX, lY = H.v2s_for(sites)
    C, T = H.golden_env(z, 'final_' if any(k.startswith('final_') for k in z.files) else 'mid_')
    dt = next(iter(sites.values())).dtype
    st_c, env_c = H.State(sites, v2s, lX, lY), H.Env(meta['chi'], dict(C), dict(T))
    st_g, env_g = H.State(H.to_dev(sites, dev), v2s, lX, lY), H.Env(meta['chi'], H.to_dev(C, dev), H.to_dev(T, dev))
    oracle = H.OracleEngine()
    p = next(iter(sites.values())).shape[0]
    g = torch.Generator().manual_seed(5)
    op1 = torch.randn(p, p, dtype=dt, generator=g)
    ops2 = [torch.randn(p, p, dtype=dt, generator=g) for _ in range(5)]
    t2 = torch.randn(p, p, p, p, dtype=dt, generator=g)
    t2s = [torch.randn(p, p, p, p, dtype=dt, generator=g) for _ in range(3)]
    for d in DIRS:
        chi1, d2, chi2 = edge_shapes(st_c, env_c, d)
        V = torch.randn(chi1, d2, chi2, dtype=dt, generator=g)
        monkeypatch.setattr(oc, '_engine', lambda: oracle)
        w_edge = oc.get_edge((0, 0), d, st_c, env_c)
        w_tm = oc.apply_TM_1sO((0, 0), d, st_c, env_c, V, op=op1)
        Ve = torch.randn(*w_edge.shape, dtype=dt, generator=g)
        w_s = oc.apply_edge((0, 0), d, st_c, env_c, Ve)
        w_c = oc.corrf_1sO1sO((0, 0), d, st_c, env_c, op1, lambda r: ops2[r], 4)
        monkeypatch.setattr(oc, '_engine', lambda: eng)
        assert H.maxrel(oc.get_edge((0, 0), d, st_g, env_g).cpu(), w_edge) < 1e-13
        assert H.maxrel(oc.apply_TM_1sO((0, 0), d, st_g, env_g, V.to(dev), op=op1.to(dev)).cpu(), w_tm) < 1e-13
        assert abs(complex(oc.apply_edge((0, 0), d, st_g, env_g, Ve.to(dev)).cpu()) - complex(w_s)) < 1e-13 * abs(complex(w_s)) + 1e-15 * float(Ve.abs().max() * w_edge.abs().max()) * Ve.numel()
        got = oc.corrf_1sO1sO((0, 0), d, st_g, env_g, op1.to(dev), lambda r: ops2[r].to(dev), 4)
        assert got.device.type == 'cuda' and H.maxrel(got.cpu(), w_c) < 1e-11, (d, got, w_c)
        # two-site operators along (MPO bond on the edge) and across (width-2 transfer matrix; down / right) the direction
        rev = (-d[0], -d[1])
        monkeypatch.setattr(oc, '_engine', lambda: oracle)
        W = torch.randn(*oc.get_edge_2((0, 0), rev, st_c, env_c).shape, dtype=dt, generator=g)
        want = [oc.apply_TM_2sO_1sChannel((0, 0), d, st_c, env_c, V, op=t2), oc.corrf_2sOH2sOH_E1((0, 0), d, st_c, env_c, t2, lambda r: t2s[r], 2),
                oc.get_edge_2((0, 0), d, st_c, env_c), oc.apply_edge((1, 0), d, st_c, env_c, W).reshape(1)]
        if d in ((0, 1), (1, 0)):
            want += [oc.apply_TM_2sO_2sChannel((0, 0), d, st_c, env_c, W, op=t2),
                     oc.corrf_2sOV2sOV_E2((0, 0), d, st_c, env_c, t2, lambda r: t2s[r], 2)]
        monkeypatch.setattr(oc, '_engine', lambda: eng)
        t2g, Wg = t2.to(dev), W.to(dev)
        have = [oc.apply_TM_2sO_1sChannel((0, 0), d, st_g, env_g, V.to(dev), op=t2g),
                oc.corrf_2sOH2sOH_E1((0, 0), d, st_g, env_g, t2g, lambda r: t2s[r].to(dev), 2),
                oc.get_edge_2((0, 0), d, st_g, env_g), oc.apply_edge((1, 0), d, st_g, env_g, Wg).reshape(1)]
        if d in ((0, 1), (1, 0)):
            have += [oc.apply_TM_2sO_2sChannel((0, 0), d, st_g, env_g, Wg, op=t2g),
                     oc.corrf_2sOV2sOV_E2((0, 0), d, st_g, env_g, t2g, lambda r: t2s[r].to(dev), 2)]
        for i, (x, y) in enumerate(zip(have, want)):
            assert x.device.type == 'cuda' and H.maxrel(x.cpu(), y) < 1e-10, (d, i, x, y)


@pytest.mark.parametrize('name', ['c4v_D2_chi8_B', 'c4v_D2_chi8_B_c128'])
def test_c4v_correlation_functions(eng, dev, name, monkeypatch):
    """ctm/one_site_c4v/corrf_c4v.py (edges, one- and two-site transfer matrices with operators, corrf_1sO1sO,
    corrf_2sOH2sOH_E1) and get_Top_spec_c4v on libctmb against the same functions with the oracle as engine (pinned against
    the unmodified reference by tests/test_transferops_cpu.py)."""
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    from peps_torch_b200.ctm.one_site_c4v import corrf_c4v as oc, transferops_c4v as ot
    from test_transferops_cpu import c4v_case
    a, chi, C, T = c4v_case(name)
    dt = a.dtype
    st_c, st_g = IPEPS_C4V(a), IPEPS_C4V(a.to(dev))
    env_c, env_g = ENV_C4V(chi, st_c), ENV_C4V(chi, st_g)
    env_c.C[env_c.keyC], env_c.T[env_c.keyT] = C.clone(), T.clone()
    env_g.C[env_g.keyC], env_g.T[env_g.keyT] = C.to(dev), T.to(dev)
    oracle = H.OracleEngine()
    p = a.shape[0]
    g = torch.Generator().manual_seed(9)
    op1 = torch.randn(p, p, dtype=dt, generator=g)
    ops1 = [torch.randn(p, p, dtype=dt, generator=g) for _ in range(5)]
    op2 = torch.randn(p, p, p, p, dtype=torch.float64, generator=g).to(dt)
    ops2 = [torch.randn(p, p, p, p, dtype=torch.float64, generator=g).to(dt) for _ in range(4)]
    V = torch.randn(chi, a.shape[1] ** 2, chi, dtype=dt, generator=g)
    W = torch.randn(chi, a.shape[1] ** 2, a.shape[1] ** 2, chi, dtype=dt, generator=g)
    monkeypatch.setattr(oc, '_engine', lambda: oracle)
    want = [oc.get_edge(st_c, env_c), oc.apply_TM_1sO(st_c, env_c, V, op=op1), oc.apply_TM_2sO(st_c, env_c, V, op=op2),
            oc.apply_edge(st_c, env_c, V).reshape(1), oc.corrf_1sO1sO(st_c, env_c, op1, lambda r: ops1[r], 4),
            oc.corrf_2sOH2sOH_E1(st_c, env_c, op2, lambda r: ops2[r], 3),
            oc.get_edge_L(st_c, env_c, l=2), oc.apply_TM_1sO_2(st_c, env_c, W, op=op2), oc.apply_edge_L(st_c, env_c, W).reshape(1),
            oc.corrf_2sOV2sOV_E2(st_c, env_c, op2, lambda r: ops2[r], 3)]
    Lw, L2w = ot.get_Top_spec_c4v(3, st_c, env_c), ot.get_Top2_spec_c4v(2, st_c, env_c)
    monkeypatch.setattr(oc, '_engine', lambda: eng)
    Vg = V.to(dev)
    got = [oc.get_edge(st_g, env_g), oc.apply_TM_1sO(st_g, env_g, Vg, op=op1.to(dev)), oc.apply_TM_2sO(st_g, env_g, Vg, op=op2.to(dev)),
           oc.apply_edge(st_g, env_g, Vg).reshape(1), oc.corrf_1sO1sO(st_g, env_g, op1.to(dev), lambda r: ops1[r].to(dev), 4),
           oc.corrf_2sOH2sOH_E1(st_g, env_g, op2.to(dev), lambda r: ops2[r].to(dev), 3),
           oc.get_edge_L(st_g, env_g, l=2), oc.apply_TM_1sO_2(st_g, env_g, W.to(dev), op=op2.to(dev)),
           oc.apply_edge_L(st_g, env_g, W.to(dev)).reshape(1),
           oc.corrf_2sOV2sOV_E2(st_g, env_g, op2.to(dev), lambda r: ops2[r].to(dev), 3)]
    for i, (x, y) in enumerate(zip(got, want)):
        assert x.device.type == 'cuda' and H.maxrel(x.cpu(), y) < 1e-10, (i, x, y)
    assert float((ot.get_Top_spec_c4v(3, st_g, env_g).cpu() - Lw).abs().max()) < 1e-10
    assert float((ot.get_Top2_spec_c4v(2, st_g, env_g).cpu().abs().sort(0)[0] - L2w.abs().sort(0)[0]).abs().max()) < 1e-9
    # entanglement spectrum of the L-leg cylinder (ring MPO of T tensors, transferops_c4v.py:119-181)
    from test_transferops_cpu import _spec
    monkeypatch.setattr(oc, '_engine', lambda: oracle)
    from peps_torch_b200.ctm.generic import corrf as gc
    monkeypatch.setattr(gc, '_engine', lambda: oracle)
    Sw = ot.get_EH_spec_Ttensor(2, 4, st_c, env_c)
    monkeypatch.setattr(gc, '_engine', lambda: eng)
    assert float((_spec(ot.get_EH_spec_Ttensor(2, 4, st_g, env_g).cpu()) - _spec(Sw)).abs().max()) < 1e-9


def test_tma_fed_gemm_layouts_and_edges(eng, dev):
    """tc_kernel_tma (tc_gemm_tma.cu): plain strided operands of large real contractions are fed by cp.async.bulk.tensor.
    All four fast-direction pairs (2-D swizzled map for k-fast, 3-D map for m-fast operands), extents that are not multiples
    of the 128 x 128 tile (out-of-bounds rows are zero-filled by the TMA unit), transposed output, strided views."""
    g = torch.Generator().manual_seed(11)
    for (M, N, K) in ((2048, 2000, 512), (1992, 2048, 1024), (4096, 4096, 256)):
        A = torch.randn(M, K, dtype=torch.float64, generator=g).to(dev)
        B = torch.randn(K, N, dtype=torch.float64, generator=g).to(dev)
        At, Bt = A.t().contiguous(), B.t().contiguous()
        want = A @ B
        scale = float(want.abs().max())
        for spec, X, Y in (('ik,kj->ij', A, B), ('ki,kj->ij', At, B), ('ik,jk->ij', A, Bt), ('ki,jk->ij', At, Bt),
                           ('ik,kj->ji', A, B), ('ki,jk->ji', At, Bt)):
            out = eng.einsum2(spec, X, Y)
            ref = want if spec.endswith('ij') else want.t()
            assert float((out - ref).abs().max()) < 1e-12 * scale, (spec, M, N, K)
    # operands that are views with a leading dimension larger than their extent (the blocks of the blocked QR, slabs)
    big = torch.randn(2304, 2304, dtype=torch.float64, generator=g).to(dev)
    A, B = big[:2048, 128:128 + 512], big[256:256 + 512, :2048]
    out = eng.einsum2('ik,kj->ij', A.contiguous(), B.contiguous())
    assert float((out - A @ B).abs().max()) < 1e-12 * float((A @ B).abs().max())


def test_kagome_density_matrices_on_gpu(eng, dev, monkeypatch):
    """trace1x1_dn_kagome / rdm2x2_dn_triangle_with_operator / rdm2x2_up_triangle_open (ctm/pess_kagome/rdm_kagome.py) on
    libctmb against the same functions with the oracle as engine (pinned against the unmodified reference and the energies of
    models/spin_half_kagome.py by tests/test_kagome_rdm_cpu.py); also at the size of config 4 (D = 3, chi = 64, p = 8)."""
    from peps_torch_b200.ctm.pess_kagome import rdm_kagome as ok
    from test_kagome_rdm_cpu import kagome_fixture
    sites, v2s, lX, lY, C, T, chi = kagome_fixture()
    cases = [(sites, v2s, lX, lY, C, T, chi)]
    a = orc.random_state_kagome(3, family='B')
    s4 = OrderedDict({(0, 0): a})
    C4, T4 = orc.init_env(s4, orc.v2s_1site, 64)
    for d in orc.DIRECTIONS:
        orc.ctm_move(d, s4, orc.v2s_1site, C4, T4, 64)
    cases.append((s4, orc.v2s_1site, 1, 1, C4, T4, 64))
    oracle = H.OracleEngine()
    g = torch.Generator().manual_seed(3)
    op = torch.randn(8, 8, dtype=torch.float64, generator=g)
    for sites, v2s, lX, lY, C, T, chi in cases:
        st_c, env_c = H.State(sites, v2s, lX, lY), H.Env(chi, dict(C), dict(T))
        st_g, env_g = H.State(H.to_dev(sites, dev), v2s, lX, lY), H.Env(chi, H.to_dev(C, dev), H.to_dev(T, dev))
        monkeypatch.setattr(ok, '_engine', lambda: oracle)
        w1 = ok.trace1x1_dn_kagome((0, 0), st_c, env_c, op)
        w2, wn = ok.rdm2x2_dn_triangle_with_operator((0, 0), st_c, env_c, op)
        w3 = ok.rdm2x2_up_triangle_open((0, 0), st_c, env_c)
        monkeypatch.setattr(ok, '_engine', lambda: eng)
        g1 = ok.trace1x1_dn_kagome((0, 0), st_g, env_g, op.to(dev))
        g2, gn = ok.rdm2x2_dn_triangle_with_operator((0, 0), st_g, env_g, op.to(dev))
        g3 = ok.rdm2x2_up_triangle_open((0, 0), st_g, env_g)
        assert abs(float(g1) - float(w1)) < 1e-11 * abs(float(w1))
        assert abs(float(g2) - float(w2)) < 1e-11 and abs(float(gn) - float(wn)) < 1e-11 * abs(float(wn))
        assert float((g3.cpu() - w3).abs().max()) < 1e-12
