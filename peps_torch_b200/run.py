"""Launcher that runs an UNMODIFIED peps-torch script with the CTM move replaced by libctmb:

    python -m peps_torch_b200.run <peps-torch>/examples/j1j2/ctmrg_j1j2.py --tiling 4SITE --bond_dim 3 --chi 48 \
           --GLOBALARGS_device cuda:0

The reference's move modules look ``ctm_MOVE`` up as a module global at call time
(ctm/generic/ctmrg.py:68, ctm/one_site_c4v/ctmrg_c4v.py:89) and the scripts hold the module object
(``from ctm.generic import ctmrg``), so rebinding the attribute is seen everywhere.
Everything else of the reference (states, models, RDMs, config, argparse) is used as it is.
"""
import importlib
import torch
import os
import runpy
import sys


def find_reference_root(script):
    """The scripts' own ``context.py`` prepends <script dir>/../.. (examples/j1j2/context.py:1-3)."""
    root = os.environ.get('PEPS_TORCH_ROOT')
    if root:
        return os.path.abspath(root)
    return os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(script)), '..', '..'))


def enable(engine_factory=None):
    """Rebind the hot-path functions of the (already importable) reference to libctmb.
    ``engine_factory`` is for host-logic tests only; the default is the CUDA engine (no fallback)."""
    from . import ctm as _pkg                                   # noqa: F401  (package import check)
    from .ctm.generic import ctmrg as ours
    from .ctm.one_site_c4v import ctmrg_c4v as ours_c4v
    from .ctm.generic import rdm as ours_rdm
    from .ctm.one_site_c4v import rdm_c4v as ours_rdm_c4v
    if engine_factory is not None:
        for m in (ours, ours_c4v, ours_rdm, ours_rdm_c4v):
            m._engine = engine_factory
    ref = importlib.import_module('ctm.generic.ctmrg')
    ref_c4v = importlib.import_module('ctm.one_site_c4v.ctmrg_c4v')

    def ctm_MOVE(direction, state, env, ctm_args=ref.cfg.ctm_args, global_args=ref.cfg.global_args,
                 verbosity=0, diagnostics=None):
        return ours.ctm_MOVE(direction, state, env, ctm_args=ctm_args, global_args=global_args,
                             verbosity=verbosity, diagnostics=diagnostics)

    def ctm_MOVE_sl(a, env, f_c2x2_decomp=None, ctm_args=ref_c4v.cfg.ctm_args, global_args=ref_c4v.cfg.global_args,
                    past_steps_data=None):
        return ours_c4v.ctm_MOVE_sl(a, env, f_c2x2_decomp, ctm_args=ctm_args, global_args=global_args,
                                    past_steps_data=past_steps_data)

    def ctm_MOVE_dl(a, env, f_c2x2_decomp=None, ctm_args=ref_c4v.cfg.ctm_args, global_args=ref_c4v.cfg.global_args):
        return ours_c4v.ctm_MOVE_dl(a, env, f_c2x2_decomp, ctm_args=ctm_args, global_args=global_args)

    # the reference's own run / run_overlap / run_dl loops stay: they look these names up at call time, build the
    # double-layer tensors themselves under ctm_force_dl and hand rank-4 sites to ctm_MOVE, which libctmb accepts
    ref.ctm_MOVE = ctm_MOVE
    def ctm_MOVE_QR_sl(a, env, ctm_args=ref_c4v.cfg.ctm_args, global_args=ref_c4v.cfg.global_args, past_steps_data=None):
        return ours_c4v.ctm_MOVE_QR_sl(a, env, ctm_args=ctm_args, global_args=global_args, past_steps_data=past_steps_data)

    ref_c4v.ctm_MOVE_sl = ctm_MOVE_sl
    ref_c4v.ctm_MOVE_dl = ctm_MOVE_dl
    ref_c4v.ctm_MOVE_QR_sl = ctm_MOVE_QR_sl
    # the plaquette density matrices behind the energies of the J1-J2 scripts (models/j1j2.py:223-247,641-679) run on
    # libctmb as well
    # Under autograd (optim_*.py: the energy is differentiated with respect to the state, through the environment) the
    # density matrices stay the reference's own torch code -- libctmb's RDM entry points are forward-only; AD is built for
    # the move (peps_torch_b200/ad.py).  Everything evaluated under torch.no_grad() (convergence
    # checks, observables) runs on libctmb.
    from .ad import needs_grad

    def _tensors(args):
        for x in args:
            for attr in ('sites', 'C', 'T'):
                d = getattr(x, attr, None)
                if isinstance(d, dict):
                    yield from d.values()

    def _dispatch(ours_fn, ref_fn):
        def f(*args, **kw):
            if needs_grad(list(_tensors(args)) + list(_tensors(kw.values()))):
                return ref_fn(*args, **kw)
            return ours_fn(*args, **kw)
        f.__name__ = getattr(ours_fn, '__name__', 'rdm')
        return f

    rdm = importlib.import_module('ctm.generic.rdm')
    ref_legacy = rdm.rdm2x2_legacy
    for name in ('rdm2x2', 'rdm2x2_legacy', 'rdm2x2_oe', 'rdm1x1', 'rdm2x1', 'rdm1x2', 'rdm1x1_dl', 'rdm2x1_dl', 'rdm1x2_dl',
                 'rdm1x1_sl', 'rdm2x1_sl', 'rdm1x2_sl'):
        ref_fn = getattr(rdm, name)
        if name == 'rdm2x2':
            # the reference's dispatching rdm2x2 needs opt_einsum (ctm/generic/rdm.py:1354-1362); its pure-torch
            # rdm2x2_legacy takes (coord, state, env, sym_pos_def, ...) and is what runs under autograd
            def ref_fn(coord, state, env, open_sites=[0, 1, 2, 3], sym_pos_def=False, **kw):
                if list(open_sites) != [0, 1, 2, 3]:
                    raise NotImplementedError("rdm2x2 with open_sites under autograd needs the reference's opt_einsum path")
                return ref_legacy(coord, state, env, sym_pos_def=sym_pos_def)
        setattr(rdm, name, _dispatch(getattr(ours_rdm, name), ref_fn))
    rdm_c4v = importlib.import_module('ctm.one_site_c4v.rdm_c4v')
    for name in ('rdm2x2_NN_lowmem_sl', 'rdm2x2_NNN_lowmem_sl', 'rdm2x2_NN_lowmem', 'rdm2x2_NNN_lowmem', 'rdm2x2',
                 'rdm1x1', 'rdm1x1_sl', 'rdm2x1', 'rdm2x1_sl', 'rdm3x1', 'rdm3x1_sl'):
        setattr(rdm_c4v, name, _dispatch(getattr(ours_rdm_c4v, name), getattr(rdm_c4v, name)))
    # the transfer-operator spectra at the tail of the scripts (ctm/generic/transferops.py:38-205): mat-vecs on libctmb
    from .ctm.generic import transferops as ours_top
    top = importlib.import_module('ctm.generic.transferops')
    if engine_factory is not None:
        from .ctm.generic import corrf as ours_corrf
        ours_corrf._engine = engine_factory
    top.get_Top_spec = ours_top.get_Top_spec
    top.get_Top_w0_spec = ours_top.get_Top_w0_spec
    top.get_EH_spec_Ttensor = ours_top.get_EH_spec_Ttensor
    # the two-point functions behind eval_corrf_* of the models (ctm/generic/corrf.py:10-104,234-277,364-650,980-1067);
    # double-layer (rank-4) sites and anything under autograd stay the reference's torch code
    from .ctm.generic import corrf as ours_cf
    cf = importlib.import_module('ctm.generic.corrf')
    for name in ('get_edge', 'get_edge_2', 'apply_edge', 'apply_TM_1sO', 'apply_TM_2sO_1sChannel', 'apply_TM_2sO_2sChannel',
                 'corrf_1sO1sO', 'corrf_2sOH2sOH_E1', 'corrf_2sOV2sOV_E2'):
        def f(coord, direction, state, env, *args, _ours=getattr(ours_cf, name), _ref=getattr(cf, name), **kw):
            extra = [x for x in list(args) + list(kw.values()) if isinstance(x, torch.Tensor)]
            single_layer = all(t.dim() == 5 for t in state.sites.values())
            if not single_layer or needs_grad(list(_tensors((state, env))) + extra):
                return _ref(coord, direction, state, env, *args, **kw)
            return _ours(coord, direction, state, env, *args, **kw)
        f.__name__ = name
        setattr(cf, name, f)
    # the same for the C4v ansatz (ctm/one_site_c4v/corrf_c4v.py, transferops_c4v.py:10-68): eval_corrf_SS / eval_corrf_DD_H
    # and the transfer-operator spectrum at the tail of ctmrg_j1j2_c4v.py
    from .ctm.one_site_c4v import corrf_c4v as ours_cf4, transferops_c4v as ours_top4
    if engine_factory is not None:
        ours_cf4._engine = engine_factory
    cf4 = importlib.import_module('ctm.one_site_c4v.corrf_c4v')
    for name in ('get_edge', 'apply_edge', 'get_edge_L', 'apply_edge_L', 'apply_TM_1sO', 'apply_TM_1sO_2', 'apply_TM_2sO',
                 'corrf_1sO1sO', 'corrf_2sOH2sOH_E1', 'corrf_2sOV2sOV_E2'):
        ref_fn = getattr(cf4, name)

        def f(state, env, *args, _ours=getattr(ours_cf4, name), _ref=ref_fn, **kw):
            extra = [x for x in args if isinstance(x, torch.Tensor)]
            if needs_grad(list(_tensors((state, env))) + extra):
                return _ref(state, env, *args, **kw)
            return _ours(state, env, *args, **kw)
        setattr(cf4, name, f)
    top4 = importlib.import_module('ctm.one_site_c4v.transferops_c4v')
    top4.get_Top_spec_c4v = ours_top4.get_Top_spec_c4v
    top4.get_Top2_spec_c4v = ours_top4.get_Top2_spec_c4v
    top4.get_EH_spec_Ttensor = ours_top4.get_EH_spec_Ttensor
    # the kagome density matrices behind energy_triangle_dn / _up and eval_obs of models/spin_half_kagome.py (config 4)
    from .ctm.pess_kagome import rdm_kagome as ours_kag
    if engine_factory is not None:
        ours_kag._engine = engine_factory
    try:
        kag = importlib.import_module('ctm.pess_kagome.rdm_kagome')
        for name in ('trace1x1_dn_kagome', 'rdm2x2_dn_triangle_with_operator', 'rdm2x2_up_triangle_open'):
            setattr(kag, name, _dispatch(getattr(ours_kag, name), getattr(kag, name)))
    except ImportError:
        pass
    # (this also removes the reference's dependence on opt_einsum for these functions: without it its 'sl' one- and
    # two-site RDMs and the rdm2x2 dispatch do not run at all, ctm/generic/rdm.py:107-112,292,343-351,560,1354-1362)
    return ref, ref_c4v


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = argv[0]
    root = find_reference_root(script)
    for p in (os.path.dirname(os.path.abspath(script)), root):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    enable()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
