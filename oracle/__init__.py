"""Parity oracle for vinum_b200 -- test infrastructure only (see vinum_oracle.py)."""
