"""CPU oracle of the spectrum-sense hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under scanner_b200/ does.
"""
from .binding import *  # noqa: F401,F403
