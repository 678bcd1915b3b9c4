#!/bin/bash
# compute-sanitizer passes over small instances of every kernel family (memcheck + racecheck on the
# shared-memory PCR, synccheck on the cluster kernels).  Slow (each kernel is instrumented): the selection below keeps
# the systems short.  Writes gpurun_out/sanitize_<tool>.log; prints one summary line per tool.
#     gpurun --timeout 900 -- 'bash tools/sanitize_gpu.sh'
mkdir -p gpurun_out
SEL='test_assembly_batched_and_bc_variants_bit_exact or test_assembly_nodal_h_many_wavenumbers_bit_exact or (test_random_systems_all_kernel_shapes and (2049 or 4097 or 63)) or test_fused_equals_staged or (test_periodic_recursion_ragged_sizes and (1040 or 4112 or 131089)) or test_fused_plans_replay or test_spike_pieces_on_stored_bands'
for tool in memcheck racecheck synccheck; do
    timeout 280 compute-sanitizer --tool $tool --error-exitcode 99 --log-file gpurun_out/sanitize_$tool.log \
        python -m pytest tests/test_gpu_1d.py tests/test_gpu_large.py -x -q -k "$SEL" > gpurun_out/sanitize_$tool.out 2>&1
    rc=$?
    echo "$tool: exit $rc; $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log 2>/dev/null) summaries; $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log | sort | uniq -c | tr '\n' ';')"
    tail -2 gpurun_out/sanitize_$tool.out
done
