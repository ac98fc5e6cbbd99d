"""Stage and import the UNMODIFIED reference (TEST / BENCH INFRASTRUCTURE ONLY).

The reference is three Python files with no build step.  ``stage()`` is the committed recipe that puts
byte-identical copies of them under the git-ignored ``oracle/_ref/`` (never under version control, see
``.gitignore``; not listed in ``.gpurunignore``, so the directory travels to the GPU box like the built
``.so``), where ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time the reference's own
``MCEvidence(...).evidence()`` on the box's host cores (``kind: "reference"``).  ``__graft_entry__.build()``
calls ``stage()`` whenever ``/root/reference`` is present (the build container); on the GPU box the staged
files are used as they arrived.

    python oracle/ref_loader.py          # stage + import check

``load()`` imports the staged (or mounted) module under the name ``reference_MCEvidence`` through the
two-line ``DistanceMetric`` shim of SURVEY.md F10 -- sklearn moved the class the reference imports at
``MCEvidence.py:37``; no line of the reference is edited.  Only ``tests/``, ``bench.py``'s CPU legs and
``__graft_entry__`` may import this module; the product package never does.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
SOURCE_DIR = os.environ.get('MCE_REFERENCE_DIR', '/root/reference')
FILES = ('MCEvidence.py', 'planck_mcevidence.py', 'examples.py')


def _sha(path):
    with open(path, 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage(source_dir=SOURCE_DIR):
    """Copy the reference's own sources, unmodified, into oracle/_ref/.  Returns {file: sha256}."""
    if not os.path.exists(os.path.join(source_dir, 'MCEvidence.py')):
        raise FileNotFoundError('reference sources not found under %s' % source_dir)
    os.makedirs(REF_DIR, exist_ok=True)
    out = {}
    for name in FILES:
        src = os.path.join(source_dir, name)
        if os.path.exists(src):
            dst = os.path.join(REF_DIR, name)
            if not os.path.exists(dst) or _sha(dst) != _sha(src):
                shutil.copyfile(src, dst)
            out[name] = _sha(dst)
    return out


def available():
    return os.path.exists(os.path.join(REF_DIR, 'MCEvidence.py')) or os.path.exists(os.path.join(SOURCE_DIR, 'MCEvidence.py'))


def reference_dir():
    if os.path.exists(os.path.join(REF_DIR, 'MCEvidence.py')):
        return REF_DIR
    if os.path.exists(os.path.join(SOURCE_DIR, 'MCEvidence.py')):
        return SOURCE_DIR
    raise FileNotFoundError('the reference is neither staged under oracle/_ref/ (python oracle/ref_loader.py) '
                            'nor mounted at %s' % SOURCE_DIR)


_mod = None


def load():
    """The reference module, unmodified, imported as ``reference_MCEvidence`` (the drop-in module of this
    repository is called ``MCEvidence`` too)."""
    global _mod
    if _mod is not None:
        return _mod
    import sklearn.metrics
    import sklearn.neighbors
    sklearn.neighbors.DistanceMetric = sklearn.metrics.DistanceMetric
    path = os.path.join(reference_dir(), 'MCEvidence.py')
    spec = importlib.util.spec_from_file_location('reference_MCEvidence', path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules['reference_MCEvidence'] = mod
    spec.loader.exec_module(mod)
    _mod = mod
    return mod


if __name__ == '__main__':
    for k, v in stage().items():
        print('%s  %s' % (v[:16], os.path.join('oracle/_ref', k)))
    m = load()
    print('imported', m.__file__, 'version', getattr(m, '__version__', '?'))
