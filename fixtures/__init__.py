"""Test / benchmark fixtures (NOT part of the product package): deterministic synthetic clouds for
the BASELINE.json configs and the reference's own inline test fixtures.  Used by tests/, bench.py,
tools/ and __graft_entry__.smoke()."""
