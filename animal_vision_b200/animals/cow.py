"""Drop-in for reference animals/cow.py."""
from .mammals import Cow  # noqa: F401
