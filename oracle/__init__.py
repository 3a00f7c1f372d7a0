"""CPU oracle (test infrastructure only; see fa_oracle.py header)."""
