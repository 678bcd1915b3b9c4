"""Prints the few numbers of a bench.py JSON line that matter while tuning a kernel."""
import json
import sys

b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st, sg = b["stages"], b["staged"]
print("step %.3f ms (one-pass kernel %.3f, frac %.3f; nets %.3f, %s refined) | staged %.3f ms: assembly %.3f + solve %.3f (frac %.3f) "
      "+ nets %.3f | fused %.3f | berr %.2e / %.2e | identical %s" % (
          b["ms_per_step"], st["one_pass_kernel_ms"], b["roofline"]["frac"], st["refine_and_repair_ms"],
          st.get("refined_systems_per_step_max_over_ranks"), sg["ms_per_step"], sg["assembly_ms"], sg["solve_ms"],
          sg["solve_roofline"]["frac"], sg["refine_and_repair_ms"], b.get("fused", {}).get("ms_per_step", float("nan")),
          b["parity"]["backward_error_max"], sg["backward_error_max"], sg["identical_to_one_pass_outputs"]))
e = b.get("e2e")
if e:
    print("e2e %.1f ms (%.1f GB/s over all GPUs, ceiling %.1f GB/s, frac %.2f) | l2-reduced %.2f ms" % (
        e["ms_per_step"], e["d2h_gbs_all_gpus"], e["ceiling_gbs"], e["frac_of_ceiling"], e["reduced_l2"]["ms_per_step"]))
for name, c in (b.get("other_configs") or {}).items():
    print(name, json.dumps(c)[:900])
