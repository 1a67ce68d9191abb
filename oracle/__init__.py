"""CPU oracle (test infrastructure only) — see oracle/sph_oracle.c."""
