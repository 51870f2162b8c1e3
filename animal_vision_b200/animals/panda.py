"""Drop-in for reference animals/panda.py."""
from .mammals import Panda  # noqa: F401
