"""Prints the few numbers of a bench.py JSON line that matter while tuning a kernel."""
import json
import sys

b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = b["stages"]
print("step %.3f ms | assembly %.3f | solve %.3f (frac %.3f) | refine %.3f (%s systems) | fused %.3f | one-pass %.3f | berr %.2e" % (
    b["ms_per_step"], st["assembly_ms"], st["solve_ms"], b["roofline"]["frac"], st["refine_ms"],
    st.get("refined_systems_per_step_max_over_ranks"), b.get("fused", {}).get("ms_per_step", float("nan")),
    b.get("one_pass", {}).get("ms_per_step", float("nan")),
    b["parity"]["backward_error_max"]))
e = b.get("e2e")
if e:
    print("e2e %.1f ms (%.1f GB/s over all GPUs, ceiling %.1f GB/s, frac %.2f) | l2-reduced %.2f ms" % (
        e["ms_per_step"], e["d2h_gbs_all_gpus"], e["ceiling_gbs"], e["frac_of_ceiling"], e["reduced_l2"]["ms_per_step"]))
for name, c in (b.get("other_configs") or {}).items():
    print(name, json.dumps(c)[:900])
