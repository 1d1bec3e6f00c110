"""srp-b200: a B200-native implementation of the draw path of kitrofimov/srp.

The product is the C library (include/srp/*.h + include/srp_b200.h, sources under
srp_b200/csrc, built by srp_b200/build.py); this package only holds the build recipe and a
ctypes mirror of the C API used by the tests and the benchmark.  Importing the package does
not load CUDA; `load_product()` does and raises if the library has not been built."""
from .host import load_product, SrpLibrary  # noqa: F401

__all__ = ["load_product", "SrpLibrary"]
