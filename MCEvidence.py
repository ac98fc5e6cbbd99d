#!/usr/bin/env python
"""Drop-in ``MCEvidence`` module and command line backed by the B200 CUDA hot path.

``from MCEvidence import MCEvidence`` and ``python MCEvidence.py <root> -k K --burn B [--cross]``
keep working as with the reference (class, module-level helpers and flags of
``MCEvidence.py:613-1473`` there); the neighbour search and reduction run in
``mcevidence_b200/libmcevidence_b200.so``.  Flags are unchanged; ``-b/--burnfrac`` and
``-t/--thinfrac`` are added as aliases of ``--burn`` / ``--thin`` because the reference README
and ``planck_mcevidence.py`` use those spellings.
"""
from __future__ import absolute_import, print_function

import logging
import os
import sys
from argparse import ArgumentParser

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mcevidence_b200 import priors as _priors                       # noqa: E402
from mcevidence_b200.chains import FORMAT, LoggingHandler, MCSamples, SamplesMIXIN, data_set   # noqa: E402,F401
from mcevidence_b200.evidence import MCEvidence                      # noqa: E402,F401
from mcevidence_b200.priors import (cosmo_params_list, extract_array, extract_dict, get_prior_volume,   # noqa: E402,F401
                                    iscosmo_param, params_info, query_yes_no)

use_getdist = False      # the reference's import probe (MCEvidence.py:50-59); getdist is never needed here

__version_info__ = ('17', '04', '2018')          # interface version of the reference this mirrors
__version__ = '-'.join(__version_info__)
__status__ = 'B200'

desc = 'Planck Chains MCEvidence. Returns the log Bayesian Evidence computed using the kth NN'
cite = '''
**
When using this code in published work, please cite the following paper: **
Heavens et. al. (2017)
Marginal Likelihoods from Monte Carlo Markov Chains
https://arxiv.org/abs/1704.03472
'''

logger = logging.getLogger(__name__)


def build_parser(prog=None):
    parser = ArgumentParser(prog=prog, add_help=True, description=desc, epilog=cite)
    parser.add_argument('root_name', help='Root filename for MCMC chains or python class filename')
    parser.add_argument('--version', action='version',
                        version='>>>   %(prog)s :: {0} version date: {1}   <<<'.format(__status__, __version__))
    parser.add_argument('-k', '--kmax', dest='kmax', default=2, type=int, help='scikit maximum K-NN ')
    parser.add_argument('-ic', '--idchain', dest='idchain', default=0, type=int,
                        help='Which chains to use - the id e.g 1 means read only *_1.txt '
                             '(default=None - use all available) ')
    parser.add_argument('-np', '--ndim', dest='ndim', default=None, type=int,
                        help='How many parameters to use (default=None - use all params) ')
    parser.add_argument('--paramsfile', dest='paramsfile', default='', type=str,
                        help='text file name containing additional parameter names to consider as '
                             'cosmological parameters')
    parser.add_argument('--burn', '--burnlen', '-b', '--burnfrac', dest='burnlen', default=0, type=float,
                        help='Burn-in length or fraction. burnlen<1 is interpreted as fraction e.g. 0.3 - 30%%')
    parser.add_argument('--thin', '--thinlen', '-t', '--thinfrac', dest='thinlen', default=0, type=float,
                        help='''Thinning fraction.
                             If 0<thinlen<1, MCMC weights are adjusted based on Poisson sampling
                             If thinlen>1, weighted thinning based on getdist algorithm
                             If thinlen<0, thinning length will be the autocorrelation length of the chain
                             ''')
    parser.add_argument('-vb', '--verbose', dest='verbose', default=1, type=int,
                        help='Increase output verbosity: 0: WARNNINGS, 1: INFO, 2: DEBUG, >2: EVERYTHING')
    parser.add_argument('-pv', '--pvolume', dest='priorvolume', default=None, type=float,
                        help='prior volume to use. If *.range exist, prior_volume estimated internally is used.')
    parser.add_argument('--allparams', help='flag to use all params and not use iscosmo_params condition',
                        action='store_true')
    parser.add_argument('--cross', action='store_true',
                        help='flag to split chain (s) to estimate cross Evidence. Otherwise auto Evidence is '
                             'calculated. The nearest neighbour of a point of one half is searched in the other half.')
    return parser


def main(argv=None):
    args = build_parser(prog=sys.argv[0]).parse_args(argv)

    if args.paramsfile != '':
        new_names = []
        with open(args.paramsfile, 'r') as fp:
            for one_line in fp:
                line = one_line.strip()
                if line != '' and line.find('#') == -1:
                    new_names.append(line)
        print('Adding additional parameter names to cosmo_params list from %s..' % args.paramsfile)
        print('adding the following names:', new_names)
        for nm in new_names:
            if nm not in _priors.cosmo_params_list:
                _priors.cosmo_params_list.append(nm)

    logging.basicConfig(level=logging.INFO, format=FORMAT)
    level = logging.DEBUG if args.verbose > 1 else (logging.INFO if args.verbose == 1 else logging.WARNING)
    for name in (__name__, 'mcevidence_b200.priors', 'mcevidence_b200.chains', 'mcevidence_b200.evidence'):
        logging.getLogger(name).setLevel(level)

    prior_volume = get_prior_volume(args, cosmo=not args.allparams)

    print()
    print('Using file: ', args.root_name)
    mce = MCEvidence(args.root_name, split=args.cross, ndim=args.ndim, priorvolume=prior_volume,
                     idchain=args.idchain, kmax=args.kmax, verbose=args.verbose, burnlen=args.burnlen,
                     thinlen=args.thinlen)
    mle = mce.evidence()
    print('* ln(B)[k] is the natural logarithm of the Baysian evidence estimated using the kth Nearest Neighbour.')
    print('')
    return mle


if __name__ == '__main__':
    main()
