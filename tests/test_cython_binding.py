"""the reference's OWN Python binding (bindings/python/edlib.pyx + cedlib.pxd, Cython) is
built against this repository's header and library instead of the bundled edlib.cpp -- only the build recipe
changes -- and the reference's own bindings/python/test.py must pass on it.
Here the library is the CPU emulation build of the engine; skipped where the reference tree is absent."""
import os
import shutil
import subprocess
import sys

import pytest

from edlib_b200._ffi import REPO

BINDING = "/root/reference/bindings/python"

SETUP = """
from setuptools import setup, Extension
from Cython.Build import cythonize
setup(name="edlib", ext_modules=cythonize([Extension("edlib", ["edlib.pyx"], include_dirs=[%(inc)r],
      libraries=["edlib_emul"], library_dirs=[%(lib)r], runtime_library_dirs=[%(lib)r], language="c++")],
      compiler_directives={"language_level": "3"}))
"""


@pytest.mark.skipif(not os.path.exists(os.path.join(BINDING, "edlib.pyx")), reason="reference binding sources not present")
def test_reference_cython_binding_over_this_library(tmp_path):
    pytest.importorskip("Cython")
    subprocess.run(["make", "-s", "-C", os.path.join(REPO, "tests", "emul")], check=True)
    for f in ("edlib.pyx", "cedlib.pxd", "test.py"):
        shutil.copy(os.path.join(BINDING, f), tmp_path / f)
    (tmp_path / "setup.py").write_text(SETUP % dict(inc=os.path.join(REPO, "include"), lib=os.path.join(REPO, "tests", "emul")))
    subprocess.run([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp_path, check=True, capture_output=True)
    out = subprocess.run([sys.executable, "test.py"], cwd=tmp_path, check=True, capture_output=True, text=True)
    assert "All tests passed!" in out.stdout, out.stdout[-400:]


@pytest.mark.gpu
def test_reference_cython_binding_over_the_cuda_library():
    """The same binding built over the PRODUCT library (`make -C oracle pybinding`, prebuilt into the git-ignored
    oracle/_ref/pybinding, which travels to the GPU box): the reference's own bindings/python/test.py on the GPU."""
    d = os.path.join(REPO, "oracle", "_ref", "pybinding")
    if not os.path.exists(os.path.join(d, "test.py")):
        pytest.skip("reference binding was not built (no /root/reference where the snapshot was taken)")
    out = subprocess.run([sys.executable, "test.py"], cwd=d, capture_output=True, text=True)
    assert out.returncode == 0 and "All tests passed!" in out.stdout, (out.stdout[-400:], out.stderr[-400:])
