"""TEST INFRASTRUCTURE ONLY -- CPU oracle of ExaChem's fused CCSD(T) path.

Importable only from tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs and the measurement tools under tools/ that compare against the reference (gpu_comparator, benzene_real).  The product package (exachem_b200) never imports this.
"""
