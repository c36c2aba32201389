"""``loops_b200.algorithms`` mirrors ``loops::algorithms`` of the reference."""
from . import spmv  # noqa: F401
