"""pytest configuration: the `gpu` marker, repo paths, oracle back ends and the product loader."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# deterministic pair-order sums in the reference build; set before libgomp loads
os.environ.setdefault("OMP_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle_built():
    from oracle import host as H
    if not os.path.exists(H.PORT_SO) or (os.path.isdir("/root/reference") and not os.path.exists(H.REF_SO)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)


@pytest.fixture(scope="session")
def H():
    """oracle.host (the fake MATLAB host + oracle back ends); builds the CPU oracles when missing."""
    _ensure_oracle_built()
    from oracle import host
    return host


@pytest.fixture(scope="session")
def port(H):
    return H.PortBackend()


@pytest.fixture(scope="session")
def ref(H):
    if not os.path.exists(H.REF_SO):
        pytest.skip("oracle/_ref/libsph_ref.so not built (needs /root/reference)")
    return H.RefBackend()


@pytest.fixture(scope="session")
def oracle_be(H):
    """The checker used by parity tests: the reference build when present, else the pinned port."""
    return H.best_backend()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


def load_product():
    """Import the package directory `sph-poiseuille-flow_b200/` (not a valid identifier) as sph_b200."""
    name = "sph_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "sph-poiseuille-flow_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    # Test infrastructure only: SPHB_CPUSIM_LIB points the ctypes mirror at the CPU functional simulator build of
    # the same sources (tests/cpusim, driven by tests/test_cpusim.py).  The product itself knows no such switch.
    sim = os.environ.get("SPHB_CPUSIM_LIB")
    if sim:
        mod.LIB_PATH = sim
        # "device" memory of the simulator is host memory: the single-process slab group keeps its exchange
        # buffers in CPU tensors there
        import torch
        slab = importlib.import_module(name + ".slab")
        group_init = slab.SlabGroup.__init__

        def init_on_cpu(self, cfg, pt, n_slabs, device="cpu", **kw):
            group_init(self, cfg, pt, n_slabs, device="cpu", **kw)

        slab.SlabGroup.__init__ = init_on_cpu
        torch.cuda.synchronize = lambda *a, **k: None
    return mod


@pytest.fixture(scope="session")
def sphb():
    return load_product()
