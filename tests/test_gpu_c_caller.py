"""The drop-in boundary from plain C: tests/c/abi_caller.c is compiled with gcc against include/mcevidence_b200.h, linked
with libmcevidence_b200.so only (no Python, no torch in that process) and must print the numbers of the Python wrapper."""
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from mcevidence_b200.synthetic import gaussian_chain

pytestmark = pytest.mark.gpu


def test_plain_c_caller_matches_python(tmp_path, ctx):
    exe = str(tmp_path / 'abi_caller')
    libdir = os.path.join(ROOT, 'mcevidence_b200')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c', 'abi_caller.c'), '-L', libdir, '-lmcevidence_b200', '-lm',
                           '-Wl,-rpath,' + libdir, '-o', exe])
    ldd = subprocess.check_output(['ldd', exe]).decode()
    assert 'python' not in ldd.lower() and 'torch' not in ldd.lower()
    chain, _ = gaussian_chain(20011, 7, seed=52, nextra=3, mean_shift=0.1)
    path = str(tmp_path / 'chain.bin')
    with open(path, 'wb') as fh:
        np.array(chain.shape, dtype=np.int64).tofile(fh)
        np.ascontiguousarray(chain).tofile(fh)
    out = subprocess.run([exe, path, '7', '4'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln.split() for ln in out.stdout.strip().splitlines()]
    s, logl, w = chain[:, 2:], -chain[:, 1], chain[:, 0]
    lsw = math.log(w.sum())
    want = ctx.evidence(s, logl, w, 4, ndim=7, log_sum_w=lsw)['logE']
    single = np.array([float(x) for x in lines[0][1:]])
    assert lines[0][0] == 'single'
    assert np.abs(single - want).max() < 1e-12      # the C program sums the weights in its own order
    assert lines[1][0] == 'group' and int(lines[1][1]) >= 1
    group = np.array([float(x) for x in lines[1][3:]])
    assert np.abs(group - want).max() < 1e-12
    n = len(chain)
    for ln, rows in zip(lines[2:5], (n // 5, n // 2, n)):
        assert ln[0] == 'sweep' and int(ln[1]) == rows
        got = np.array([float(x) for x in ln[2:]])
        if rows == n:
            assert np.abs(got - want).max() < 1e-12
        assert np.all(np.isfinite(got))
    for ln, i in zip(lines[5:8], range(3)):
        assert ln[0] == 'batch' and int(ln[1]) == i
        sub = ctx.evidence(s[i::3], logl[i::3], w[i::3], 4, ndim=7, log_sum_w=math.log(w[i::3].sum()))['logE']
        assert np.abs(np.array([float(x) for x in ln[2:]]) - sub).max() < 1e-12
