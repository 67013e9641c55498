"""CPU oracle package - TEST INFRASTRUCTURE ONLY (see gl_oracle.py header)."""
