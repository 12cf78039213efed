"""The committed branch-coverage table of the taumol evaluators (profiles/taumol_branch_coverage.{md,json}) is what
tools/taumol_coverage.py produces from the current kernel code, and says what DESIGN.md claims: no reference golden reaches a
water-vapour term that contributes, the CORK cross-check does."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("taumol_coverage", os.path.join(ROOT, "tools", "taumol_coverage.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_committed_coverage_table_is_current_and_says_what_the_docs_claim():
    cov = _tool().collect()
    committed = json.load(open(os.path.join(ROOT, "profiles", "taumol_branch_coverage.json")))
    assert cov == committed, "re-run python tools/taumol_coverage.py"
    for which in ("lw", "sw"):
        for key, row in cov[which].items():
            for name in ("self>0", "foreign>0"):
                if name in row:
                    assert "R" not in row[name], (which, key, name)      # every reference golden is a dry column
                    assert "X" in row[name], (which, key, name)          # ... the CORK cross-check reaches them all
    # every band has layers in both regions in every set of states
    assert all(row.get("region") == "RXS" for which in cov for row in cov[which].values())
    # the three-point stencils (specparm < 0.125 / > 0.875) are only ever entered in the lower atmosphere of binary bands
    assert "s0<.125" in cov["lw"]["band 3 lower"] and "s0>.875" in cov["lw"]["band 3 lower"]
