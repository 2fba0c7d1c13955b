"""CPU oracle (test infrastructure only) -- see oracle/tha4_oracle.py."""
