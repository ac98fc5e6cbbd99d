"""Host-side chain container: loading, burn-in, thinning, split (reference MCEvidence.py:99-606)."""
import numpy as np
import pytest

from mcevidence_b200.chains import MCSamples, read_chain_file


def make(method, **kw):
    kw.setdefault('log_level', 30)
    return MCSamples(method, **kw)


def test_partitions_match_reference(golden, tmp_path):
    """Same rows and weights in s1 / s2 as the reference built from the same input and RNG seed."""
    from collections import namedtuple
    g = golden
    sv = namedtuple('split_var', 'split frac shuffle')
    np.random.seed(g.np_seed)
    kw = {k: g.ctor[k] for k in ('burnlen', 'thinlen') if k in g.ctor}
    ms = make(g.method(tmp_path), csplit=sv(g.ctor.get('split', False), g.ctor.get('s1frac', 0.5), True), **kw)
    np.testing.assert_array_equal(ms.data['s1'].samples[:, 0], g.raw['s1_col0'])
    np.testing.assert_array_equal(ms.data['s1'].weights, g.raw['s1_weights'])
    if 's2_col0' in g.raw.files:
        np.testing.assert_array_equal(ms.data['s2'].samples[:, 0], g.raw['s2_col0'])
        np.testing.assert_array_equal(ms.data['s2'].weights, g.raw['s2_weights'])
    else:
        assert ms.data['s2'].samples is None
    assert list(g.raw['nsample']) == [ms.get_shape(n)[0] for n in (['s1', 's2'] if g.ctor.get('split') else ['s1'])]


def test_array_input_ignores_burn_and_thin():
    """Reference quirk kept for parity: keywords are not forwarded for array input (MCEvidence.py:151)."""
    c = np.random.default_rng(0).random((100, 5))
    ms = make([c], burnlen=0.5, thinlen=0.5)
    assert ms.get_shape()[0] == 100


def test_arrays_sign_and_columns():
    c = np.random.default_rng(1).random((50, 6))
    ms = make((c, c[:10]))
    s, lnp, w = ms.arrays('s1')
    assert s.shape == (60, 4)
    np.testing.assert_array_equal(lnp, -np.concatenate((c[:, 1], c[:10, 1])))
    np.testing.assert_array_equal(w, np.concatenate((c[:, 0], c[:10, 0])))
    assert ms.nparamMC == 4 and ms.names == ['p0', 'p1', 'p2', 'p3'] and ms.labels[0] == 'x_0'
    d = make({'a': c, 'b': c})
    assert d.get_shape()[0] == 100
    with pytest.raises(TypeError):
        make(3.5)


def test_file_inputs(tmp_path):
    rng = np.random.default_rng(2)
    root = str(tmp_path / 'ch')
    chains = [rng.random((30 + i, 5)) for i in range(3)]
    for i, c in enumerate(chains):
        np.savetxt('%s_%d.txt' % (root, i + 1), c, fmt='%.10e')
    rounded = [np.loadtxt('%s_%d.txt' % (root, i + 1)) for i in range(3)]
    for i in range(3):
        np.testing.assert_array_equal(read_chain_file('%s_%d.txt' % (root, i + 1)), rounded[i])
    assert make(root).get_shape()[0] == 30 + 31 + 32                      # root_?.txt
    assert make(root, idchain=2).get_shape()[0] == 31                     # root_2.txt
    assert make(root + '_1.txt').get_shape()[0] == 30                     # a file
    assert make([root + '_1.txt', root + '_3.txt']).get_shape()[0] == 62  # list of files
    assert make(root + '_*.txt').get_shape()[0] == 93                     # wildcard
    np.testing.assert_array_equal(make(root).data['s1'].samples, np.concatenate(rounded)[:, 2:])
    # burn-in: fraction per chain, or rows per chain (a float row count is accepted)
    assert make(root, burnlen=0.5).get_shape()[0] == 15 + 16 + 16
    assert make(root, burnlen=10.0).get_shape()[0] == 20 + 21 + 22
    with pytest.raises(IOError):
        make(str(tmp_path / 'nothing'))


def thin_indices_loop(weights, factor):
    """Plain transcription of the getdist integer-weight algorithm, the definition thin_indices must meet."""
    w = [int(x) for x in weights]
    out, tot, i = [], 0, 0
    mult = w[0]
    while i < len(w):
        if mult + tot < factor:
            tot += mult
            i += 1
            if i < len(w):
                mult = w[i]
        else:
            out.append(i)
            if mult == factor - tot:
                i += 1
                if i < len(w):
                    mult = w[i]
            else:
                mult -= factor - tot
            tot = 0
    return np.array(out)


@pytest.mark.parametrize('factor', [2, 3, 7, 50])
def test_integer_thinning(factor):
    rng = np.random.default_rng(factor)
    w = rng.integers(1, 12, size=400).astype(float)
    c = np.column_stack((w, rng.random(400), rng.random((400, 3))))
    ms = make([c])
    ix, nw = ms.thin_indices(factor, weights=w.copy())
    if factor >= w.max():
        _, want = np.unique(np.cumsum(w.astype(int)) // factor, return_index=True)
    else:
        want = thin_indices_loop(w, factor)
    np.testing.assert_array_equal(ix, want)
    np.testing.assert_array_equal(nw, w[ix])
    with pytest.raises(ValueError):
        ms.thin_indices(2, weights=w + 0.25)
    # non-integer weights fall back to the heaviest sample of each bin
    nw2, ix2 = ms.get_thin_index(4, w + 0.25)
    assert len(ix2) == 100 and np.all(np.diff(ix2) > 0)
    for b, i in enumerate(ix2[1:-1], start=1):
        assert (w + 0.25)[i] == (w + 0.25)[np.digitize(np.arange(400), np.linspace(-1, 400, 101)) == b + 1].max()


def test_poisson_thin_stream_and_partition_thin():
    rng = np.random.default_rng(3)
    w = rng.integers(1, 9, size=300).astype(float)
    c = np.column_stack((w, rng.random(300), rng.random((300, 2))))
    ms = make([c])
    np.random.seed(9)
    ix, nw = ms.poisson_thin(0.4, weights=w)
    np.random.seed(9)
    ref = np.array([float(np.random.poisson(x)) for x in w * 0.4])    # the reference's scalar loop (:427)
    np.testing.assert_array_equal(ix, np.where(ref > 0)[0])
    np.testing.assert_array_equal(nw, ref[ref > 0])
    np.random.seed(9)
    ms.thin(0.4, name='s1')
    assert ms.get_shape()[0] == len(ix)
    np.testing.assert_array_equal(ms.data['s1'].adjusted_weights, nw)
    assert ms.thin(1) is None
    with pytest.raises(NotImplementedError):
        ms.get_thin_index(-1, w)


def test_importance_sample_and_burn_partition():
    c = np.random.default_rng(4).random((40, 4))
    ms = make([c])
    ms.importance_sample(lambda s: s[:, 0], name='s1')
    np.testing.assert_allclose(ms.data['s1'].adjusted_weights, c[:, 0] * np.exp(-c[:, 2]))
    np.testing.assert_array_equal(ms.data['s1'].weights, c[:, 0])
    ms.removeBurn(0.25, name='s1')
    assert ms.get_shape()[0] == 30
    assert ms.removeBurn(7) == 7


def test_native_chain_reader_matches_loadtxt(tmp_path):
    """mce_read_chain_text (multi-threaded from_chars, strtod fallback) returns exactly what np.loadtxt returns."""
    from mcevidence_b200 import _lib
    rng = np.random.default_rng(6)
    x = rng.normal(size=(20000, 9)) * 10.0 ** rng.integers(-20, 20, size=(20000, 9))
    f = str(tmp_path / 'wide.txt')
    np.savetxt(f, x, fmt='%.17g')
    for nt in (1, 3, 0):
        np.testing.assert_array_equal(_lib.read_chain_text(f, nt), x)
    g = str(tmp_path / 'messy.txt')
    with open(g, 'w') as fh:
        fh.write('# weight loglike p1\n1 2.5 3e-3\n\n  4 5 6   # trailing comment\n7\t8\t-9.0E+1\r\n   \n10 11 12')
    want = np.loadtxt(g)
    np.testing.assert_array_equal(read_chain_file(g), want)
    assert read_chain_file(g).flags.c_contiguous and read_chain_file(g).shape == (4, 3)
    r = str(tmp_path / 'ragged.txt')
    with open(r, 'w') as fh:
        fh.write('1 2 3\n4 5\n')
    with pytest.raises(_lib.MCEError):
        _lib.read_chain_text(r)
    with pytest.raises(ValueError):          # numpy's own error for what the fast path rejects
        read_chain_file(r)
    one = str(tmp_path / 'one.txt')
    with open(one, 'w') as fh:
        fh.write('1 2 3 4\n')
    assert read_chain_file(one).shape == (1, 4)
    # spellings the fast conversion hands to strtod or has to treat like Python's float(): leading '+', inf / nan,
    # overflow, underflow to a denormal and to zero, bare leading / trailing decimal points, signed zero
    odd = str(tmp_path / 'odd.txt')
    with open(odd, 'w') as fh:
        fh.write('+1.5 -2e3\t3.25E-2 inf\r\n 4 5 6 nan # c\n7.0e+00 8 9 -inf\n 1e-320 1e400 1e-400 -0.0\n.5 5. 1E5 -.25e-1\n')
    got, want = _lib.read_chain_text(odd, 2), np.loadtxt(odd)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(np.signbit(got), np.signbit(want))
    bad = str(tmp_path / 'bad.txt')
    with open(bad, 'w') as fh:
        fh.write('1 2x 3\n')
    with pytest.raises(_lib.MCEError):       # trailing garbage in a token is an error, as in np.loadtxt
        _lib.read_chain_text(bad)
