"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product package).

ctypes loader for oracle/mvsv_oracle.c.  Allowed importers: tests/,
__graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
"""
from .loader import *  # noqa: F401,F403
