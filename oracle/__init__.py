"""CPU oracle: TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs;
never from boxlcd_b200/ (the product path has no CPU fallback)."""
