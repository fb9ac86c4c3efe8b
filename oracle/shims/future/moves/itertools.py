"""Stub of future.moves.itertools."""
from itertools import *  # noqa: F401,F403
