"""CPU oracle for the DandD sketch-and-count hot path.  TEST INFRASTRUCTURE ONLY -- see the header
of dandd_oracle.c.  Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; the product package (dandd_b200) never imports this."""
