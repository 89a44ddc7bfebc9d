"""CPU oracle (test infrastructure only; see pp_oracle.py header)."""
