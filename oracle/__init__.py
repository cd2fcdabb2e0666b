"""CPU oracle for jubjub_b200 -- TEST INFRASTRUCTURE ONLY (see jj_oracle.h)."""
