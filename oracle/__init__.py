"""CPU oracle for the commit-phase hot path.  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the product."""
