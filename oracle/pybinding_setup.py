"""Test infrastructure: build recipe of the reference's OWN Cython binding (bindings/python/edlib.pyx + cedlib.pxd,
copied next to this script's working directory by `make -C oracle pybinding`) against this repository's header and
PRODUCT library instead of the bundled edlib.cpp -- the recipe INTEGRATION.md section 2 describes.  Outputs stay under
the git-ignored oracle/_ref/."""
import os

from Cython.Build import cythonize
from setuptools import Extension, setup

inc, lib, name = os.environ["EB_INC"], os.environ["EB_LIBDIR"], os.environ["EB_LIBNAME"]
setup(name="edlib", ext_modules=cythonize([Extension("edlib", ["edlib.pyx"], include_dirs=[inc], libraries=[name],
      library_dirs=[lib], runtime_library_dirs=[lib], language="c++")], compiler_directives={"language_level": "3"}))
