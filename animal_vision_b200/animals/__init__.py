"""Species plugins, one module per species as in the reference's animals/ package."""
from .animal import Animal  # noqa: F401
from .mammals import *  # noqa: F401,F403
from .mammals import MAMMALS  # noqa: F401
from .cat import Cat  # noqa: F401
from .honeybee import HoneyBee  # noqa: F401
