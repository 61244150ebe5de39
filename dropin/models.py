"""Drop this file next to the reference scripts as ``models.py`` (all four model directories):
every name any script imports from ``models`` resolves to the B200 implementation."""
from multimodalbrainsurvival_b200.models import *  # noqa: F401,F403
