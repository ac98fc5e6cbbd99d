"""mcevidence_b200: the MCEvidence kNN-evidence hot path on B200 (sm_100a).

``MCEvidence`` keeps the reference's constructor / ``evidence()`` API; the covariance, whitening,
exact k-th nearest-neighbour search and the log-volume reduction run in hand-written CUDA kernels
behind the C ABI of ``include/mcevidence_b200.h`` (``libmcevidence_b200.so``, loaded with ctypes).
There is no CPU fallback: importing the package is cheap, the first compute call loads the library.
"""
from .chains import MCSamples, SamplesMIXIN, data_set          # noqa: F401
from .evidence import MCEvidence                                # noqa: F401
from .priors import cosmo_params_list, get_prior_volume, iscosmo_param, params_info   # noqa: F401

__all__ = ['MCEvidence', 'MCSamples', 'SamplesMIXIN', 'data_set', 'params_info', 'get_prior_volume',
           'iscosmo_param', 'cosmo_params_list']
