"""Drop-in for reference animals/rabbit.py."""
from .mammals import Rabbit  # noqa: F401
