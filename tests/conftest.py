"""pytest configuration: registers the `gpu` marker and puts tests/ on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The product library and the example binary are built in-tree by `make` and normally travel with the repository
    snapshot; on a checkout without them, build them once (nvcc cross-compiles without a GPU).  The tests themselves
    never fall back to anything: a missing library is an error."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    need = [os.path.join(root, "odr_audioenc_b200", "libtoolame_b200.so"), os.path.join(root, "examples", "dabenc"),
            os.path.join(root, "oracle", "_build", "libmp2_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-s", "-C", root], check=False, stdout=subprocess.DEVNULL)
        subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle"), "all"], check=False, stdout=subprocess.DEVNULL)
