"""The RDM-distance convergence criterion of the C4v scripts (examples/j1j2/ctmrg_j1j2_c4v.py:101-129) as
peps_torch_b200.env.ctmrg_conv_rdm2x1.  Known answer: the unmodified script on the config-1 state
(tests/golden/config1_instate.json, chi 16, j2 0.3) prints the distances inf, inf, 5.108689210750049e-08, 7.243128631977052e-09
and stops at the fourth move.  CPU: host logic with the oracle as engine; the GPU suite runs the same check on libctmb."""
import os
import pytest
import torch
import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF_DISTS = [float('inf'), float('inf'), 5.108689210750049e-08, 7.243128631977052e-09]


def run_config1(dev, lag):
    from peps_torch_b200.ipeps import read_ipeps, IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V, init_env_c4v, ctmrg_conv_rdm2x1
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    from peps_torch_b200.config import CTMARGS
    st = IPEPS_C4V(read_ipeps(os.path.join(GOLD, 'config1_instate.json'), device=dev).sites[(0, 0)])
    env = ENV_C4V(16, st)
    init_env_c4v(st, env)
    args = CTMARGS()
    args.ctm_max_iter, args.ctm_conv_tol = 50, 1.0e-8
    env, hist, t_ctm, t_obs = ctmrg_c4v.run(st, env, conv_check=lambda s, e, h, ctm_args=args: ctmrg_conv_rdm2x1(s, e, h, ctm_args, lag=lag),
                                            ctm_args=args)
    return hist


def check(hist, lag):
    log = hist['log']
    assert len(log) == len(REF_DISTS)                  # lag 1: one move more, the last distance still pending
    assert log[0] == log[1] == float('inf')
    for got, want in zip(log[2:], REF_DISTS[2:]):
        assert abs(got - want) < 1e-12 + 1e-4 * want, (log, REF_DISTS)     # distances of 1e-8 between RDMs accurate to 1e-13
    assert len(hist.get('pending', [])) == lag


@pytest.mark.parametrize('lag', [0, 1])
def test_conv_rdm2x1_history_matches_the_reference_script(monkeypatch, lag):
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v, rdm_c4v
    eng = H.OracleEngine()
    monkeypatch.setattr(ctmrg_c4v, '_engine', lambda: eng)
    monkeypatch.setattr(rdm_c4v, '_engine', lambda: eng)
    check(run_config1('cpu', lag), lag)
