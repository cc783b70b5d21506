"""Drop-in for the density matrices of ctm/one_site_c4v/rdm_c4v.py used by the C4v J1-J2 energy and observables
(models/j1j2.py:641-679, eval_obs): rdm2x2_NN_lowmem(_sl) :1117-1202, rdm2x2_NNN_lowmem(_sl) :1286-1371, rdm2x2 :1446-1546,
rdm1x1(_sl) :168-392, rdm2x1(_sl) :394-665, rdm3x1(_sl) :667-1011 (the J3 term of energy_1x1_lowmem, models/j1j2.py:671-677).
The reference rotates ONE enlarged corner; here the single (C, T) pair is rotated into the eight tensors of a generic
1x1-cell environment (env_c4v.py:25-45 vs env.py:57-77) and libctmb contracts the generic 2x2 network, tracing the
sites that are not kept (closed corners)."""


def _engine():
    from ...engine import default_engine
    return default_engine()


def _generic_tensors(state, env):
    a = next(iter(state.sites.values()))
    C_, T = env.C[env.keyC], env.T[env.keyT]
    # order of engine.C_KEYS = (-1,-1),(1,-1),(1,1),(-1,1) and engine.T_KEYS = (0,-1),(-1,0),(0,1),(1,0)
    Cs = [C_, C_, C_, C_.t().contiguous()]
    Tud = T.permute(1, 2, 0).contiguous()
    Ts = [Tud, T, T.permute(2, 0, 1).contiguous(), Tud]
    return [(a, Cs, Ts)] * 4, env.chi


def _rdm(state, env, open_sites, sym_pos_def):
    t4, chi = _generic_tensors(state, env)
    return _engine().rdm2x2_sites(t4, chi, open_sites, sym_pos_def)


def rdm2x2_NN_lowmem_sl(state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    r""":return: nearest-neighbour density matrix :math:`s_0s_1;s'_0s'_1` of the 2x2 plaquette (rdm_c4v.py:1160-1202)"""
    return _rdm(state, env, (0, 1), sym_pos_def)


def rdm2x2_NNN_lowmem_sl(state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    r""":return: next-nearest-neighbour (diagonal) density matrix :math:`s_0s_3;s'_0s'_3` (rdm_c4v.py:1329-1371)"""
    return _rdm(state, env, (0, 3), sym_pos_def)


rdm2x2_NN_lowmem = rdm2x2_NN_lowmem_sl          # the double-layer variants compute the same object
rdm2x2_NNN_lowmem = rdm2x2_NNN_lowmem_sl


def rdm2x2(state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    r""":return: 4-site density matrix :math:`s_0s_1s_2s_3;s'_0s'_1s'_2s'_3` (rdm_c4v.py:1446-1546)"""
    return _rdm(state, env, (0, 1, 2, 3), sym_pos_def)


def rdm1x1_sl(state, env, sym_pos_def=False, verbosity=0):
    r""":return: 1-site density matrix :math:`s;s'` (rdm_c4v.py:266-392)"""
    t4, chi = _generic_tensors(state, env)
    return _engine().rdm_small_sites('1x1', t4[:1], chi, sym_pos_def)


def rdm2x1_sl(state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    r""":return: 2-site density matrix :math:`s_0s_1;s'_0s'_1` of a nearest-neighbour pair (rdm_c4v.py:530-665)"""
    t4, chi = _generic_tensors(state, env)
    return _engine().rdm_small_sites('2x1', t4[:2], chi, sym_pos_def)


rdm1x1 = rdm1x1_sl
rdm2x1 = rdm2x1_sl


def rdm3x1_sl(state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    r""":return: 2-site density matrix :math:`s_0s_1;s'_0s'_1` of the two END sites of a row of three, the centre site traced
    (rdm_c4v.py:829-1011).  Built like the correlation functions (corrf_c4v.py): the left boundary C--T--C takes three
    transfer matrices -- the first and the third with the physical indices of both layers left open -- and is closed with
    the right boundary; every step is one libctmb chain and no double-layer tensor is formed."""
    from ... import ad
    eng = _engine()
    a = next(iter(state.sites.values()))
    C_, T = env.C[env.keyC], env.T[env.keyT]
    E = ad.contract(eng, 'bcd,ce->bde', ad.contract(eng, 'ab,acd->bcd', C_, T), C_)                  # left boundary
    E = ad.sl_chain(eng, 'axu,xyz,@uydr,zed->arepq', (T, E, T), a, phys='pq')                          # site 0, open
    E = ad.sl_chain(eng, 'axu,xyzpq,@uydr,zed->arepq', (T, E.contiguous(), T), a)                     # centre, traced
    E = ad.sl_chain(eng, 'axu,xyzpq,@uydr,zed->arepqvw', (T, E.contiguous(), T), a, phys='vw')         # site 1, open
    R = ad.contract(eng, 'xty,tz->xyz', ad.contract(eng, 'xc,tcy->xty', C_, T), C_)                   # right boundary
    rho = ad.contract(eng, 'arepqvw,are->pvqw', E.contiguous(), R)
    return eng.sym_pos_def(rho, sym_pos_def)


rdm3x1 = rdm3x1_sl
