"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see gpy_oracle.py header; parity unpinned)."""
