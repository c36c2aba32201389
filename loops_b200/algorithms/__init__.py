"""``loops_b200.algorithms`` mirrors ``loops::algorithms`` of the reference."""
from . import spmv  # noqa: F401
from . import spmm  # noqa: F401
