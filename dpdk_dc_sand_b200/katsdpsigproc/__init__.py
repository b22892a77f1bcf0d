"""Minimal ``katsdpsigproc`` stand-in: just the parts of ``accel`` the beamformer operators and their
tests use (reference dependency ``katsdpsigproc==1.2``, not vendored in the reference tree), implemented
over torch CUDA tensors and streams instead of PyCUDA.  See ``accel.py``."""
__dcbf_shim__ = True
__version__ = "1.2+dcbf"
