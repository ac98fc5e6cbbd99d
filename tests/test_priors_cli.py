"""Prior-volume files and the command-line surface (reference MCEvidence.py:1173-1473)."""
import os

import numpy as np
import pytest

import MCEvidence as M
from mcevidence_b200 import priors
from mcevidence_b200.synthetic import write_cosmomc_root


def test_ranges_file(tmp_path):
    root = str(tmp_path / 'base')
    with open(root + '.ranges', 'w') as fh:
        fh.write('omegabh2  0.005 0.1\nomegach2 0.001 0.99\ntheta 0.5 10\ntau 0.01 0.8\n'
                 'nnu 3.046 3.046\ncalPlanck 0.9 1.1\nH0* N N\nlogA 2 4\n')
    cosmo = priors.params_info(root, cosmo=True)
    assert cosmo['name'] == ['omegabh2', 'omegach2', 'theta', 'tau', 'logA']
    assert cosmo['ndim'] == 5 and cosmo['nr_of_params'] == 5
    np.testing.assert_allclose(cosmo['volume'], 0.095 * 0.989 * 9.5 * 0.79 * 2.0)
    everything = priors.params_info(root, cosmo=False)
    assert everything['ndim'] == 7 and np.isinf(everything['volume'])     # unbounded H0* -> infinite volume
    assert priors.iscosmo_param('ns') and not priors.iscosmo_param('calPlanck')

    class Args:
        root_name, verbose, ndim, priorvolume = root, 0, 99, 5.0
    a = Args()
    assert np.isclose(priors.get_prior_volume(a, cosmo=True), cosmo['volume'])
    assert a.ndim == 5                         # -np is overwritten
    with pytest.raises(Exception):
        priors.params_info(str(tmp_path / 'missing'))


def test_montepython_log_param(tmp_path):
    d = tmp_path / 'mp'
    d.mkdir()
    (d / 'log.param').write_text(
        "#data.parameters['skip'] = [1, 0, 2, 0.1, 1, 'cosmo']\n"
        "data.parameters['omega_b'] = [2.2, 1.8, 3.0, 0.03, 0.01, 'cosmo']\n"
        "data.parameters['A_cib'] = [60, 0, 200, 7, 1, 'nuisance']\n"
        "data.parameters['H0'] = [0, None, None, 0, 1, 'derived']\n"
        "data.cosmo_arguments['k_pivot'] = 0.05\n")
    allp = priors.params_info(str(d))
    assert allp['name'] == ['omega_b', 'A_cib'] and np.isclose(allp['volume'], 1.2 * 200)
    assert priors.params_info(str(d), cosmo=True)['name'] == ['omega_b']
    (d / 'log.param').write_text("data.parameters['x'] = [0, None, 1, 0, 1, 'cosmo']\n")
    with pytest.raises(Exception, match='Unbounded'):
        priors.params_info(str(d))


def test_cli_flags_unchanged_plus_aliases():
    p = M.build_parser('MCEvidence.py')
    a = p.parse_args(['root'])
    assert (a.kmax, a.idchain, a.ndim, a.burnlen, a.thinlen, a.verbose, a.priorvolume) == (2, 0, None, 0, 0, 1, None)
    assert not a.allparams and not a.cross and a.paramsfile == ''
    a = p.parse_args(['root', '-k', '4', '-ic', '2', '-np', '6', '--burn', '0.3', '--thin', '0.5', '-vb', '0',
                      '-pv', '2.5', '--allparams', '--cross', '--paramsfile', 'x.txt'])
    assert (a.kmax, a.idchain, a.ndim, a.burnlen, a.thinlen, a.verbose, a.priorvolume) == (4, 2, 6, 0.3, 0.5, 0, 2.5)
    assert a.allparams and a.cross
    assert p.parse_args(['r', '--burnlen', '100', '--thinlen', '3']).burnlen == 100
    # aliases used by the reference README / planck_mcevidence.py
    assert p.parse_args(['r', '-k', '2', '-b', '0.3']).burnlen == 0.3
    assert p.parse_args(['r', '--burnfrac', '0.2', '-t', '0.4']).thinlen == 0.4
    with pytest.raises(SystemExit):
        p.parse_args(['r', '--version'])


def test_module_surface():
    for name in ('MCEvidence', 'MCSamples', 'SamplesMIXIN', 'data_set', 'params_info', 'get_prior_volume',
                 'iscosmo_param', 'cosmo_params_list', '__version__', 'extract_array', 'extract_dict', 'query_yes_no',
                 'LoggingHandler', 'use_getdist'):
        assert hasattr(M, name)
    import inspect
    sig = inspect.signature(M.MCEvidence.__init__)
    ref_order = ['self', 'method', 'ischain', 'isfunc', 'thinlen', 'burnlen', 'split', 's1frac', 'shuffle', 'ndim',
                 'kmax', 'priorvolume', 'debug', 'nsample', 'covtype', 'nbatch', 'brange', 'bscale', 'verbose', 'args']
    assert list(sig.parameters)[:len(ref_order)] == ref_order           # MCEvidence.py:614-624
    assert sig.parameters['kmax'].default == 5 and sig.parameters['covtype'].default == 'single'
    ev = inspect.signature(M.MCEvidence.evidence)
    assert list(ev.parameters) == ['self', 'verbose', 'rand', 'info', 'covtype', 'profile', 'pvolume', 'pos_lnp',
                                   'nproc', 'prewhiten']               # MCEvidence.py:950-952
    assert ev.parameters['covtype'].default == 'all' and ev.parameters['nproc'].default == -1


def test_synthetic_cosmomc_root(tmp_path):
    root = str(tmp_path / 'planck' / 'base_synth')
    names, vol = write_cosmomc_root(root, nchains=2, nrows=200)
    assert os.path.exists(root + '_2.txt') and os.path.exists(root + '.ranges')
    info = priors.params_info(root, cosmo=True)
    assert info['ndim'] == 6 and np.isclose(info['volume'], vol)
    assert priors.params_info(root, cosmo=False)['ndim'] == len(names)


def test_grid_driver_discovery_and_flags(tmp_path):
    """planck_mcevidence.py: the reference's flags, and model discovery in table order on a miniature grid."""
    import planck_mcevidence as P
    from mcevidence_b200.synthetic import write_cosmomc_grid
    a = P.build_parser().parse_args(['-k', '3', '-nc', '4', '-nd', '2', '-nm', '5', '-b', '0.3', '-t', '0.5',
                                     '-o', 'out', '--N', 'xx', '-v', '0'])
    assert (a.kmax, a.nchain, a.ndata, a.nmodel, a.burnfrac, a.thinfrac, a.outdir, a.name, a.verbose) == \
        (3, 4, 2, 5, 0.3, 0.5, 'out', 'xx', 0)
    d = P.build_parser().parse_args([])
    assert (d.kmax, d.nchain, d.outdir, d.name) == (2, 0, 'planck_mce_fullGrid_R2_H0Reiss2016', 'mce')
    assert P.build_parser().parse_args(['--burnin', '0.2', '--thinfrac', '0.1']).burnfrac == 0.2
    assert len(P.DATASETS) == 31 and len(P.MODELS) == 22 and P.MODELS[0] == 'base'
    grid = write_cosmomc_grid(str(tmp_path / 'grid'), nrows=50)
    paths, names = P.avail_model_list(str(tmp_path / 'grid'), 'plikHM_TT_lowTEB')
    assert names == ['base', 'base_omegak', 'base_mnu']                  # the reference's model order
    assert paths[0] == grid[('plikHM_TT_lowTEB', 'base')][0]
    assert P.avail_model_list(str(tmp_path / 'grid'), 'plikHM_TT_lowTEB_post_BAO', nmax=2)[1] == ['base', 'base_omegak']
    assert P.avail_model_list(str(tmp_path / 'grid'), 'nothing_here')[0] == []
    assert P.h0_column(paths[0]) == 9                                     # 6 cosmological + 3 nuisance columns, then H0*
    assert priors.params_info(paths[1], cosmo=True)['ndim'] == 7
