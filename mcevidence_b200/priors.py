"""Parameter selection and prior volume from CosmoMC ``.ranges`` / MontePython ``log.param`` files.

Fresh implementation of the behaviour of ``params_info`` / ``get_prior_volume`` /
``iscosmo_param`` (reference ``MCEvidence.py:1173-1339``).  Stays on the host: it only
produces ``ndim`` and the prior volume that enter the estimator as ``log PV``.
"""
from __future__ import annotations

import glob
import logging

import numpy as np

logger = logging.getLogger(__name__)

#: names treated as cosmological (not nuisance) parameters -- the reference's list (MCEvidence.py:84-86)
cosmo_params_list = ['omegabh2', 'omegach2', 'theta', 'tau', 'omegak', 'mnu', 'meffsterile', 'w', 'wa',
                     'nnu', 'yhe', 'alpha1', 'deltazrei', 'Alens', 'Alensf', 'fdm', 'logA', 'ns', 'nrun',
                     'nrunrun', 'r', 'nt', 'ntrun', 'Aphiphi']


def iscosmo_param(p, cosmo_params=None):
    """True if ``p`` is a cosmological parameter; ``cosmo_params`` extends the list (``MCEvidence.py:1195-1202``)."""
    if cosmo_params is not None:
        cosmo_params_list.extend(cosmo_params)
    return p in cosmo_params_list


def _parse_bound(tok, sign):
    """CosmoMC writes ``N`` for an unbounded side."""
    return sign * np.inf if tok in ('N', 'n') else float(tok)


def read_ranges(fname):
    """``name min max`` rows of a CosmoMC ``.ranges`` file -> list of ``(name, min, max)``."""
    rows = []
    with open(fname, 'r') as fh:
        for line in fh:
            tok = line.split('#')[0].split()
            if len(tok) >= 3:
                rows.append((tok[0], _parse_bound(tok[1], -1.0), _parse_bound(tok[2], +1.0)))
    return rows


def _parse_logparam_line(line):
    """``data.parameters['name'] = [mean, min, max, sigma, scale, 'role']`` -> ``(name, values)``."""
    lhs, rhs = line.split('=', 1)
    name = lhs.split('[')[-1].strip().rstrip(']').strip().strip('"').strip("'")
    vals = []
    for tok in rhs.strip().lstrip('[').rstrip().rstrip(']').split(','):
        tok = tok.strip().strip('"').strip("'")
        try:
            vals.append(int(tok))
        except ValueError:
            try:
                vals.append(float(tok))
            except ValueError:
                vals.append(tok)
    return name, vals


def extract_array(line):
    """Right-hand side of a MontePython ``name = [a, b, 'c']`` line as a list of int / float / str items
    (``MCEvidence.py:1173-1185``)."""
    rhs = line.split('=')[-1].strip()
    items = [tok.strip().strip('"').strip("'") for tok in rhs.strip(']').lstrip('[').split(',')]
    out = []
    for tok in items:
        for cast in (int, float):
            try:
                out.append(cast(tok))
                break
            except ValueError:
                continue
        else:
            out.append(tok)
    return out


def extract_dict(line):
    """``data.parameters['name'] = [...]`` -> ``(name, list)`` (``MCEvidence.py:1188-1193``)."""
    lhs = line.split('=')[0].strip()
    name = lhs.split('[')[-1].strip(']').strip('"').strip("'")
    return name, extract_array(line)


def query_yes_no(question, default='yes'):
    """Interactive yes/no prompt (``MCEvidence.py:1275-1309``): True for yes, False for no; an empty answer takes
    ``default`` ('yes', 'no' or None = an answer is required)."""
    answers = {'yes': True, 'y': True, 'ye': True, 'no': False, 'n': False}
    prompts = {None: ' [y/n] ', 'yes': ' [Y/n] ', 'no': ' [y/N] '}
    if default not in prompts:
        raise ValueError("invalid default answer: '%s'" % default)
    while True:
        print(question + prompts[default], end='')
        choice = input().lower()
        if default is not None and choice == '':
            return answers[default]
        if choice in answers:
            return answers[choice]
        print("Please respond with 'yes' or 'no' (or 'y' or 'n').")


def params_info(fname, cosmo=False, volumes={}):
    """Names, ranges and prior volume of the varied parameters (``MCEvidence.py:1204-1272``).

    A parameter counts if its range is not degenerate (``not isclose(max, min)``) and, with
    ``cosmo=True``, if it is cosmological (``iscosmo_param`` for CosmoMC, role ``'cosmo'`` for
    MontePython; ``'derived'`` parameters never count).
    """
    par = {'name': [], 'min': [], 'max': [], 'range': []}
    nr = 0
    if glob.glob('{}*.ranges'.format(fname)):
        logger.info('getting params info from COSMOMC file %s.ranges' % fname)
        for name, pmin, pmax in read_ranges(fname + '.ranges'):
            pcond = iscosmo_param(name) if cosmo else True
            if not np.isclose(pmax, pmin) and pcond:
                par['name'].append(name)
                par['min'].append(pmin)
                par['max'].append(pmax)
                par['range'].append(np.abs(pmax - pmin))
                nr += 1
    elif glob.glob('{}/log.param'.format(fname)):
        logger.info('getting params info from montepython log.params file')
        with open('{}/log.param'.format(fname), 'r') as fh:
            for line in fh:
                if '#' in line or 'data.parameters' not in line:
                    continue
                name, arr = _parse_logparam_line(line)
                pcond = arr[5] == 'cosmo' if cosmo else True
                if pcond and not arr[5] == 'derived':
                    nr += 1
                    if arr[1] == 'None' or arr[2] == 'None':
                        raise Exception('Unbounded priors are not supported - please specify priors')
                    vmin, vmax = float(arr[1]), float(arr[2])
                    par['name'].append(name)
                    par['min'].append(vmin)
                    par['max'].append(vmax)
                    par['range'].append(vmax - vmin)
    else:
        raise Exception('Could not read parameter volume from COSMOMC .ranges file or montepython log.param file')
    par['str'] = ','.join(par['name'])
    par['ndim'] = len(par['name'])
    par['nr_of_params'] = nr
    par['volume'] = np.array(par['range']).prod()
    return par


def get_prior_volume(args, **kwargs):
    """Prior volume for the CLI; also overwrites ``args.ndim`` (``MCEvidence.py:1312-1339``).

    As in the reference a missing ``.ranges`` / ``log.param`` is a hard error, which makes the
    ``-np`` and ``-pv`` flags ineffective on the command line.
    """
    par = params_info(args.root_name, **kwargs)
    if args.verbose > 1:
        print(par)
    args.ndim = par['ndim']
    logger.info('getting prior volume using cosmomc *.ranges or montepython log.param outputs')
    logger.info('prior_volume=%s' % par['volume'])
    logger.info('Number of params to use: ndim=%s' % par['ndim'])
    return par['volume']
