"""Prints the few numbers of a bench.py JSON line that matter while tuning a kernel."""
import json
import sys

b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = b["stages"]
print("step %.3f ms | assembly %.3f | solve %.3f (frac %.3f) | refine %.3f (%s systems) | fused %.3f | one-pass %.3f | berr %.2e" % (
    b["ms_per_step"], st["assembly_ms"], st["solve_ms"], b["roofline"]["frac"], st["refine_ms"],
    st.get("refined_systems_per_step"), b.get("fused", {}).get("ms_per_step", float("nan")),
    b.get("one_pass", {}).get("ms_per_step", float("nan")),
    b["parity"]["backward_error_max"]))
