# profiling aid: item timelines of the step kernels (MG_TRACE variant) for c3 and c2
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
export MESHGEN_LIB=$PWD/reinforcementlearning4meshgeneration_b200/lib/variants/trace.so
timeout 300 python tests/trace_items.py c3 gpurun_out/trace_c3 4 > gpurun_out/trace_c3.log 2>&1; tail -5 gpurun_out/trace_c3.log
timeout 300 python tests/trace_items.py c2 gpurun_out/trace_c2 4 > gpurun_out/trace_c2.log 2>&1; tail -5 gpurun_out/trace_c2.log
timeout 300 python tests/trace_items.py c1 gpurun_out/trace_c1 4 > gpurun_out/trace_c1.log 2>&1; tail -5 gpurun_out/trace_c1.log
# (the .bin files stay in gpurun_out for offline analysis)
unset MESHGEN_LIB
