"""TEST INFRASTRUCTURE ONLY -- CPU oracle of ExaChem's fused CCSD(T) path.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (exachem_b200) never imports this.
"""
