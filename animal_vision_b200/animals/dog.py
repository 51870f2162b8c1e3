"""Drop-in for reference animals/dog.py."""
from .mammals import Dog  # noqa: F401
