"""Stand-in for ``katsdpsigproc.pytest_plugin`` (reference: beamformer/unit_test/conftest.py:41):
provides the ``context`` and ``command_queue`` fixtures and skips device tests when no CUDA device exists."""
import pytest

from . import accel


@pytest.fixture
def context():
    try:
        return accel.create_some_context(interactive=False, device_filter=lambda d: d.is_cuda)
    except RuntimeError as exc:
        pytest.skip(str(exc))


@pytest.fixture
def command_queue(context):
    return context.create_command_queue()
