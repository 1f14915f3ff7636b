"""CPU oracle -- test infrastructure only (see oracle/oracle.py, oracle/deepim_oracle.c)."""
