#!/usr/bin/env python
"""Packaging of the B200 drop-in: ``pip install .`` builds ``libmcevidence_b200.so`` for sm_100a with the
Makefile of ``mcevidence_b200/csrc`` (nvcc required; no GPU is needed to build) and installs the package, the
``MCEvidence`` / ``planck_mcevidence`` drop-in modules and the ``MCEvidence.py`` script -- the same entry points
the reference's setup.py installs."""
import os
import re
import subprocess

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


def find_version():
    text = open(os.path.join(HERE, 'MCEvidence.py')).read()
    m = re.search(r"^__version_info__ = \(([^)]*)\)", text, re.M)
    if not m:
        raise RuntimeError('Unable to find the version tuple in MCEvidence.py')
    day, month, year = [p.strip().strip("'\"") for p in m.group(1).split(',')]
    return '%s.%s.%s' % (year, month, day)


class BuildWithLibrary(build_py):
    """Compile the CUDA library in-tree before the Python files are collected."""

    def run(self):
        subprocess.check_call(['make', '-C', os.path.join(HERE, 'mcevidence_b200', 'csrc'), '-j8'])
        build_py.run(self)


setup(name='mcevidence-b200',
      version=find_version(),
      description='MCEvidence (Bayesian evidence from MCMC chains) with the kNN hot path on NVIDIA B200',
      url='https://github.com/yabebalFantaye/MCEvidence',
      packages=['mcevidence_b200'],
      py_modules=['MCEvidence', 'planck_mcevidence'],
      scripts=['MCEvidence.py'],
      package_data={'mcevidence_b200': ['libmcevidence_b200.so', 'planck_grid_shapes.json']},
      cmdclass={'build_py': BuildWithLibrary},
      install_requires=['numpy'],
      python_requires='>=3.8')
