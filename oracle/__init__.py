"""CPU oracle package (TEST INFRASTRUCTURE ONLY -- see oracle/nbglm_oracle.c header)."""
