"""CPU oracle (TEST INFRASTRUCTURE): ctypes loader for oracle/libace_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  See the header of ace_oracle.c for the parity status ("unpinned" beyond the
reference's known-answer tests).
"""
from .loader import Oracle, build, lib_path  # noqa: F401
