"""Abstract names the reference type-annotates with (``katsdpsigproc.abc``)."""
from .accel import AbstractCommandQueue, AbstractContext  # noqa: F401
