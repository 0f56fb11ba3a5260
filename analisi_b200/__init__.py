"""analisi_b200 -- B200-native g(r,t) (Gofrt) for rikigigi/analisi.

The product is ``libagofrt.so`` (hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/agofrt.h``) plus the C++ host classes that mirror the reference's
Trajectory / CalculateMultiThread / BlockAverage API.  This python package only holds the build
recipe, a ctypes binding of the C ABI used by the tests and ``bench.py``, and the synthetic
trajectory generator.  There is no CPU fallback: everything here raises if the CUDA library is
missing.
"""
__version__ = "0.1.0"
