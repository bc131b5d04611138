"""python setup.py build_ext --inplace   (from this directory; needs the built ../../poreseq_b200/libporeseq_b200.so)"""
import os

import numpy
from Cython.Build import cythonize
from setuptools import Extension, setup

ROOT = os.environ.get("PORESEQ_B200_ROOT") or os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
LIBDIR = os.path.join(ROOT, "poreseq_b200")
setup(name="poreseqcpp_b200",
      ext_modules=cythonize([Extension("poreseqcpp_b200", ["poreseqcpp_b200.pyx"],
                                       include_dirs=[os.path.join(ROOT, "include"), numpy.get_include()],
                                       libraries=["poreseq_b200"], library_dirs=[LIBDIR],
                                       runtime_library_dirs=[LIBDIR])], quiet=True))
