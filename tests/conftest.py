import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_cpu_namespace():
    """Tests run identical model code on both devices: the NumPy oracle is registered as the
    Device.CPU operator namespace (test infrastructure only — the package ships GPU ops only)."""
    import oracle
    oracle.register_host()
    yield


@pytest.fixture(params=[pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def device_lib(request):
    """Binds the device library for a test: the host-emulated build of the kernels on the
    GPU-less container ("emu", small shapes), the real sm_100a library on a B200 ("cuda")."""
    import helpers
    if request.param == "emu":
        helpers.use_emulated_library()
    else:
        helpers.use_cuda_library()
    return request.param
