"""CPU oracles (test infrastructure only).  See jump_oracle.py and ps_oracle.c headers."""
