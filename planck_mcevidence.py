#!/usr/bin/env python
"""Evidence tables for a CosmoMC chain grid (Planck ``COM_CosmoParams_fullGrid_R2.00`` layout) on GPUs.

Drop-in for the reference's grid driver (``planck_mcevidence.py`` there): same flags
(``-k -nc -nd -nm -b -t -o --N -v``), same outputs -- one ``psql`` table and one csv per data set
(columns ``AllChains | chain1..n, Mean_lnE_k1, Err_lnE_k1, delta_lnE_k1, PriorVol, ndim, N_read, N_used``),
the concatenated ``mce_planck_fullgrid.txt``, ``delta_lnE_all.txt`` and the two pickles.
What changed: the reference splits data sets over MPI ranks (``planck_mcevidence.py:149-160, 306-308``)
and gathers dicts (``:436``); here data sets are dealt over the GPUs of the box -- one process per GPU
under ``torchrun`` (or a single process / single GPU) -- and every evidence runs in the CUDA library.

    python planck_mcevidence.py --rootdir COM_CosmoParams_fullGrid_R2.00 -k 2 -b 0.3
    python -m torch.distributed.run --nproc-per-node 8 planck_mcevidence.py --rootdir ... -nc 4

Layout expected: ``<rootdir>/<model>/<data_dir>/<model>_<data>_<i>.txt`` with ``<model>_<data>.ranges``
next to the chains.  New flags: ``--rootdir`` (the reference hard-codes it, ``:173``).  As in the reference
(``:344``) every evidence carries the H0 = 73.24 +- 1.74 importance weight (``:56-58``; there it needs a getdist
object, here the ``H0*`` column is located through ``<root>.paramnames``): by default it is applied to every root that
has a ``.paramnames`` file, ``--h0prior`` makes a missing file an error and ``--no-h0prior`` switches the weight off (the
default output directory then drops the reference's ``_H0Reiss2016`` suffix).
"""
from __future__ import absolute_import, print_function

import glob
import logging
import os
import pickle
import sys
from argparse import ArgumentParser

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mcevidence_b200.evidence import MCEvidence          # noqa: E402
from mcevidence_b200.priors import params_info           # noqa: E402

logger = logging.getLogger(__name__)

#: data sets and models of the reference's analysis, in its table order (planck_mcevidence.py:184, 192-196)
DATASETS = ['plikHM_TT_lowTEB', 'plikHM_TT_lowTEB_post_BAO', 'plikHM_TT_lowTEB_post_lensing',
            'plikHM_TT_lowTEB_post_H070p6', 'plikHM_TT_lowTEB_post_JLA', 'plikHM_TT_lowTEB_post_zre6p5',
            'plikHM_TT_lowTEB_post_BAO_H070p6_JLA', 'plikHM_TT_lowTEB_post_lensing_BAO_H070p6_JLA',
            'plikHM_TT_lowTEB_BAO', 'plikHM_TT_lowTEB_BAO_post_lensing', 'plikHM_TT_lowTEB_BAO_post_H070p6',
            'plikHM_TT_lowTEB_BAO_post_H070p6_JLA', 'plikHM_TT_lowTEB_lensing', 'plikHM_TT_lowTEB_lensing_post_BAO',
            'plikHM_TT_lowTEB_lensing_post_zre6p5', 'plikHM_TT_lowTEB_lensing_post_BAO_H070p6_JLA',
            'plikHM_TT_tau07plikHM_TT_lowTEB_lensing_BAO', 'plikHM_TT_lowTEB_lensing_BAO_post_H070p6',
            'plikHM_TT_lowTEB_lensing_BAO_post_H070p6_JLA', 'plikHM_TTTEEE_lowTEB', 'plikHM_TTTEEE_lowTEB_post_BAO',
            'plikHM_TTTEEE_lowTEB_post_lensing', 'plikHM_TTTEEE_lowTEB_post_H070p6', 'plikHM_TTTEEE_lowTEB_post_JLA',
            'plikHM_TTTEEE_lowTEB_post_zre6p5', 'plikHM_TTTEEE_lowTEB_post_BAO_H070p6_JLA',
            'plikHM_TTTEEE_lowTEB_post_lensing_BAO_H070p6_JLA', 'plikHM_TTTEEE_lowl_lensing',
            'plikHM_TTTEEE_lowl_lensing_post_BAO', 'plikHM_TTTEEE_lowl_lensing_post_BAO_H070p6_JLA',
            'plikHM_TTTEEE_lowTEB_lensing']
MODELS = ['base', 'base_omegak', 'base_Alens', 'base_Alensf', 'base_nnu', 'base_mnu', 'base_nrun', 'base_r', 'base_w',
          'base_alpha1', 'base_Aphiphi', 'base_yhe', 'base_mnu_Alens', 'base_mnu_omegak', 'base_mnu_w', 'base_nnu_mnu',
          'base_nnu_r', 'base_nrun_r', 'base_nnu_yhe', 'base_w_wa', 'base_nnu_meffsterile', 'base_nnu_meffsterile_r']


def build_parser():
    p = ArgumentParser(description='Planck Chains MCEvidence.')
    p.add_argument('-k', '--kmax', dest='kmax', default=2, type=int, help='scikit maximum K-NN ')
    p.add_argument('-nc', '--nchain', dest='nchain', default=0, type=int,
                   help='How many chains to use (default=None - use all available) ')
    p.add_argument('-nd', '--ndata', dest='ndata', default=0, type=int,
                   help='How many data cases to use (default=None - use all available) ')
    p.add_argument('-nm', '--nmodel', dest='nmodel', default=0, type=int,
                   help='How many model cases to use (default=None - use all chains) ')
    p.add_argument('-b', '--burnfrac', '--burnin', '--remove', dest='burnfrac', default=0, type=float,
                   help='Burn-in fraction')
    p.add_argument('-t', '--thin', '--thinfrac', dest='thinfrac', default=0, type=float, help='Thinning fraction')
    p.add_argument('-o', '--out', '--outdir', dest='outdir', default='planck_mce_fullGrid_R2_H0Reiss2016',
                   help='Output directory name (default: planck_mce_fullGrid_R2_H0Reiss2016, as in the reference, or '
                        'planck_mce_fullGrid_R2 with --no-h0prior)')
    p.add_argument('--N', '--name', dest='name', default='mce', help='base name for output files')
    p.add_argument('-v', '--verbose', dest='verbose', default=1, type=int, help='increase output verbosity')
    p.add_argument('--rootdir', default='COM_CosmoParams_fullGrid_R2.00', help='root of the CosmoMC chain grid')
    p.add_argument('--datasets', nargs='*', default=None, help='data-set names (default: the Planck 2015 list)')
    p.add_argument('--h0prior', dest='h0prior', action='store_const', const='on', default='auto',
                   help='importance-weight with H0 = 73.24 +- 1.74 as the reference always does (planck_mcevidence.py:344); '
                        'needs <root>.paramnames.  Default: applied to every root that has a .paramnames file with H0')
    p.add_argument('--no-h0prior', dest='h0prior', action='store_const', const='off', help='un-weighted evidences')
    return p


def avail_model_list(rootdir, dd, nmax=0, sorter=MODELS):
    """Chain roots and model names available for data set ``dd``, in table order (planck_mcevidence.py:211-230)."""
    roots = [x[:-len('_1.txt')] for x in glob.glob('{0}/*/*/*_{1}_1.txt'.format(rootdir, dd))]
    pairs = [(r, os.path.relpath(r, rootdir).split(os.sep)[0]) for r in roots]
    rank = {m: i for i, m in enumerate(sorter or [])}
    pairs.sort(key=lambda rm: (rank.get(rm[1], len(rank)), rm[1]))
    if nmax > 0:
        pairs = pairs[:nmax]
    return [r for r, _ in pairs], [m for _, m in pairs]


def h0_column(root, required=True):
    """Index (within the parameter columns) of ``H0*`` according to ``<root>.paramnames`` (None when the file is
    missing and the weight was not asked for explicitly)."""
    fn = root + '.paramnames'
    if not os.path.exists(fn):
        if not required:
            return None
        raise IOError('--h0prior needs %s' % fn)
    names = [line.split()[0] for line in open(fn) if line.strip()]
    for i, nm in enumerate(names):
        if nm.rstrip('*') == 'H0':
            return i
    if not required:
        return None
    raise ValueError('no H0 parameter in %s' % fn)


def h0_gauss_lnp(col):
    """-log of the H0 = 73.24 +- 1.74 importance weight (Riess et al. 2016; planck_mcevidence.py:56-58) on sample column
    ``col``; the column index comes from .paramnames and is checked against the samples."""
    def func(s):
        if col >= s.shape[1]:
            raise ValueError('H0 is parameter %d of .paramnames but the chain has only %d parameter columns' % (col, s.shape[1]))
        return 0.5 * ((s[:, col] - 73.24) / 1.74) ** 2
    return func


def analyse_dataset(args, dd, chain_exts, mce_cols):
    """One table: evidence of every model for data set ``dd`` (planck_mcevidence.py:306-404)."""
    import pandas as pd
    paths, names = avail_model_list(args.rootdir, dd, nmax=args.nmodel)
    mce = np.full((len(paths), len(mce_cols)), np.nan)
    info_cols = {k: [] for k in ('PriorVol', 'ndim', 'N_read', 'N_used')}
    vol_norm = 1.0
    for imm, (mm, fname) in enumerate(zip(names, paths)):
        if not glob.glob(fname + '*.txt'):
            print('*** not available: ', fname)
            for k, v in (('N_read', ''), ('N_used', ''), ('PriorVol', 0), ('ndim', 0)):
                info_cols[k].append(v)
            continue
        par = params_info(fname, cosmo=True)
        if mm == 'base':
            vol_norm = par['volume']                  # prior volumes are relative to the base model
        prior_volume = par['volume'] / vol_norm
        isfunc = None
        if args.h0prior != 'off':
            col = h0_column(fname, required=args.h0prior == 'on')
            if col is None:
                print('*** no H0 column known for %s (no .paramnames): evidence NOT H0-weighted' % fname)
            else:
                isfunc = h0_gauss_lnp(col)
        nc_read, nc_use = '', ''
        for icc, cext in enumerate(chain_exts):
            e, info = MCEvidence(fname + cext, ndim=par['ndim'], isfunc=isfunc, priorvolume=prior_volume,
                                 kmax=args.kmax, verbose=args.verbose, burnlen=args.burnfrac,
                                 thinlen=args.thinfrac).evidence(info=True, pos_lnp=False)
            mce[imm, icc] = e[0]
            nc_read += '%s,' % info['Nsamples_read']
            nc_use += '%s,' % info['Nsamples']
        info_cols['PriorVol'].append(prior_volume)
        info_cols['ndim'].append(par['ndim'])
        info_cols['N_read'].append(nc_read)
        info_cols['N_used'].append(nc_use)
    if len(paths) == 0 or np.all(np.isnan(mce)):
        return None
    df = pd.DataFrame(mce, index=names, columns=mce_cols)
    df_mean = df.mean(axis=1)
    if args.nchain > 0:
        df_std = df.std(axis=1)
        df['Mean_lnE_k1'] = df_mean
        df['Err_lnE_k1'] = df_std / np.sqrt(args.nchain * 1.0)
    df['delta_lnE_k1'] = df_mean - df_mean.max()
    for k, v in info_cols.items():
        df[k] = v
    return df


def main(argv=None):
    import pandas as pd
    from tabulate import tabulate
    args = build_parser().parse_args(argv)
    assert args.kmax >= 2, 'kmax must be >=2'
    if args.h0prior == 'off' and args.outdir == 'planck_mce_fullGrid_R2_H0Reiss2016':
        args.outdir = 'planck_mce_fullGrid_R2'      # the reference's directory name promises the H0 weight
    logging.basicConfig(level=[logging.WARNING, logging.INFO][min(args.verbose, 1)] if args.verbose < 2 else logging.DEBUG)

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))))
    if rank == 0:
        print(args)

    data_list = list(args.datasets) if args.datasets else list(DATASETS)
    if args.ndata > 0:
        data_list = data_list[:args.ndata]
    if args.nchain == 0:
        mce_cols, chain_exts, sub = ['AllChains'], [''], 'AllChains'
    else:
        chain_exts = ['_%s.txt' % x for x in range(1, args.nchain + 1)]
        mce_cols, sub = ['chain%s' % k for k in range(1, args.nchain + 1)], 'SingleChains'
    outdir = '{}/{}/'.format(args.outdir, sub)
    outdir_data = '{}/csv'.format(outdir)
    os.makedirs(outdir_data, exist_ok=True)
    fout_txt = '%s/%s_{}.txt' % (outdir, args.name)
    fout_df = '%s/%s_{}.csv' % (outdir_data, args.name)

    all_df = {}
    for ig, dd in enumerate(data_list):
        if ig % world != rank:                        # data sets dealt round-robin over the GPUs
            continue
        df = analyse_dataset(args, dd, chain_exts, mce_cols)
        if df is None:
            continue
        all_df[dd] = df['delta_lnE_k1']
        logging.info('--------------- {}---------'.format(dd))
        table = tabulate(df, headers='keys', tablefmt='psql', floatfmt='.2f', numalign='left')
        if args.verbose > 0:
            print(table)
        with open(fout_txt.format(dd), 'w') as fh:
            if ig == 0:
                fh.write('\n############## RootDirectory={} ########\n\n'.format(args.rootdir))
            fh.write('\n************ {} ************\n'.format(dd))
            fh.write(table)
            fh.write('\n')
        df.to_csv(fout_df.format(dd))

    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, all_df)
        all_df = {k: v for d in gathered for k, v in d.items()}
    if rank == 0:
        fmain = '{}/mce_planck_fullgrid.txt'.format(outdir)
        print('all outputs being written to ', fmain)
        with open(fmain, 'w') as out:
            for dd in data_list:
                fin = fout_txt.format(dd)
                if os.path.exists(fin):
                    out.write(open(fin).read())
                    os.remove(fin)
        pickle.dump(all_df, open('{0}/delta_lnE_all_dict.pkl'.format(outdir_data), 'wb'))
        big_df = pd.DataFrame(index=MODELS)
        for dd in data_list:                           # columns in data-set order
            if dd in all_df:
                big_df[dd] = all_df[dd]
        pickle.dump(big_df, open('{0}/delta_lnE_all_df.pkl'.format(outdir_data), 'wb'))
        with open('{0}/delta_lnE_all.txt'.format(outdir), 'w') as fh:
            fh.write('\n############## RootDirectory={} ########\n\n'.format(args.rootdir))
            short = []
            for ik, k in enumerate(big_df.keys()):
                fh.write('# C{}={} \n'.format(ik, k))
                short.append('C%s' % ik)
            big_df.columns = short
            fh.write(tabulate(big_df, headers='keys', tablefmt='psql', floatfmt='.2f', numalign='left'))
            fh.write('\n')
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return all_df


if __name__ == '__main__':
    main()
