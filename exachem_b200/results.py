"""Results section of a CCSD(T) run, as the reference's driver writes it (row f4 of SURVEY.md 8f).

`ccsd_t_results` reproduces the keys and arithmetic of exachem/cc/ccsd_t/ccsd_t.cpp:265-271 (energies),
:285-292 (`output."CCSD(T)"."[T]Energies"/"(T)Energies"`) and :332-343 (`performance`), so that the reference's
CI comparator (ci/scripts/compare_results.py:197-236) runs unchanged on a file written by `write_json_data`.
"""
from __future__ import annotations

import json


def ccsd_t_results(energy1: float, energy2: float, hf_energy: float, corr_energy: float, total_t_time: float,
                   work_time: float, total_num_ops: float) -> dict:
    """energy1 = E[T], energy2 = E(T) (already summed over ranks, ccsd_t.cpp:262-263); work_time = the ranks'
    average work time (ccsd_t.cpp:330)."""
    return {
        "[T]Energies": {"correction": energy1, "correlation": corr_energy + energy1,
                        "total": hf_energy + corr_energy + energy1},
        "(T)Energies": {"correction": energy2, "correlation": corr_energy + energy2,
                        "total": hf_energy + corr_energy + energy2},
        "performance": {"total_time": total_t_time, "gflops": total_num_ops / (total_t_time * 1e9),
                        "total_num_ops": float(total_num_ops), "load_imbalance": 1.0 - work_time / total_t_time},
    }


def write_json_data(path: str, input_options: dict, scf_final_energy: float, ccsd_correlation: float,
                    ccsd_t: dict) -> dict:
    """the subset of the reference's results file the comparator reads: input.SCF.conve, input.CC.threshold,
    output.SCF.final_energy, output.CCSD.final_energy.correlation, output."CCSD(T)"."""
    doc = {"input": input_options,
           "output": {"SCF": {"final_energy": scf_final_energy},
                      "CCSD": {"final_energy": {"correlation": ccsd_correlation,
                                                "total": scf_final_energy + ccsd_correlation}},
                      "CCSD(T)": ccsd_t}}
    with open(path, "w") as f:
        json.dump(doc, f, indent=2)
    return doc
