"""Test infrastructure: CPU oracle of the HAMT hot path (see hamt_oracle.py header)."""
