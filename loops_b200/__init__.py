"""loops-b200: Blackwell-native load-balanced SpMV schedules behind the
gunrock/loops API (layout views, ``schedule::setup``, ``algorithms::spmv``).

Python here is only the host-side mirror used by tests and ``bench.py``:
ctypes over the C ABI of ``include/loopsb.h``; torch supplies device memory,
streams and ``torch.distributed``. All compute is in ``libloopsb200.so``.
"""
from . import _lib  # noqa: F401
from . import layout  # noqa: F401
from .container import bcsr_t, coo_t, csc_t, csr_t, dia_t, ell_t  # noqa: F401
from . import algorithms  # noqa: F401
from . import generate  # noqa: F401

__version__ = "0.1.0"
