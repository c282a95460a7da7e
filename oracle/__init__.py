"""ORACLE -- test infrastructure only (see oracle/golden.hpp).  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product."""
