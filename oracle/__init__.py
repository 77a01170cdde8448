"""CPU oracle (test infrastructure only -- see oracle/d2_cpu.py and oracle/sfod_oracle.c headers)."""
