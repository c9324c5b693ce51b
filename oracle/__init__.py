"""CPU oracle + reference-build recipes: TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header)."""
