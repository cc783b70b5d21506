"""Density matrices behind the kagome energies (ctm/pess_kagome/rdm_kagome.py of peps-torch: trace1x1_dn_kagome :312-462,
rdm2x2_up_triangle_open :1008-1130, rdm2x2_dn_triangle_with_operator :1132-1284) -- what models/spin_half_kagome.py calls
for energy_triangle_dn / energy_triangle_up / eval_obs (BASELINE config 4).

A kagome iPESS / iPEPS site carries the three spins of a down triangle in one physical index of dimension p = 2^3 (ket index
= 4 s0 + 2 s1 + s2, ipeps/ipess_kagome.py:62-82).  The reference opens single spins inside its enlarged corners; here the
plaquette and one-site matrices come from libctmb with the WHOLE physical leg of the chosen sites open (ctmb_rdm2x2 with an
open_sites mask, ctmb_rdm_small) and the spins that are not wanted are traced on the resulting 8^k x 8^k tensor -- the same
numbers, one generic kernel path.
"""
import torch


def _engine():
    from ...engine import default_engine
    return default_engine()


def _op_matrix(op):
    if op.dim() == 6 and len(set(op.size())) == 1:
        return op.reshape(op.size(0) ** 3, op.size(0) ** 3)
    assert op.dim() == 2, "Invalid operator"
    return op


def _real(x):
    return x.real if x.is_complex() else x


def trace1x1_dn_kagome(coord, state, env, op, verbosity=0, force_cpu=False):
    r""":math:`Tr\{\rho_{1x1,ABC}\, op\}` (unnormalised) for an operator on the three spins of the down triangle at ``coord``
    (rdm_kagome.py:312-462); ``op`` as a :math:`p \times p` matrix or a rank-6 tensor ``[2]*6``."""
    raw = _engine().rdm_small('1x1', coord, state, env, raw=True)           # [ket, bra]
    return torch.einsum('ij,ji', raw, _op_matrix(op).to(raw.dtype))


def rdm2x2_dn_triangle_with_operator(coord, state, env, op, force_cpu=False, verbosity=0, **kwargs):
    r"""Normalised expectation value of ``op`` on the down triangle of the upper-left site of the 2x2 patch at ``coord`` and
    the norm of the patch (rdm_kagome.py:1132-1284) -> ``(value, norm)``."""
    raw = _engine().rdm2x2(coord, state, env, open_sites=(0,), raw=True)   # [ket, bra] of site s0, the other three traced
    norm = _real(torch.einsum('ii', raw))
    return torch.einsum('ij,ji', raw, _op_matrix(op).to(raw.dtype)) / norm, norm


def rdm2x2_up_triangle_open(coord, state, env, sym_pos_def=False, force_cpu=False, verbosity=0, **kwargs):
    r"""Reduced density matrix of the three spins of the "up" triangle inside the 2x2 patch at ``coord``
    (rdm_kagome.py:1008-1130): spin 1 of the site at ``coord+(1,0)``, spin 2 of ``coord+(0,1)``, spin 0 of ``coord+(1,1)``;
    rank 6, :math:`s_0 s_1 s_2; s'_0 s'_1 s'_2`."""
    eng = _engine()
    raw = eng.rdm2x2(coord, state, env, open_sites=(1, 2, 3), raw=True)     # [j k l ; J K L], each of dimension 8
    r = raw.reshape([2] * 18)
    # ket spins (a b c) of the site to the right, (d e f) of the site below, (g h i) of the diagonal site; bras in capitals.
    # Kept: b, f, g; everything else is traced.  Index order of the result as the reference's: checked element-wise against
    # rdm2x2_up_triangle_open of the unmodified reference by tests/test_kagome_rdm_cpu.py
    rdm = torch.einsum('abcdefghi' + 'aBcdeFGhi' + '->' + 'bfg' + 'BFG', r)
    return eng.sym_pos_def(rdm.contiguous(), sym_pos_def)
