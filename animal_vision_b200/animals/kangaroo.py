"""Drop-in for reference animals/kangaroo.py."""
from .mammals import Kangaroo  # noqa: F401
