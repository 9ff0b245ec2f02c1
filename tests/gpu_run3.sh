timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1500 bash tests/ab_variants.sh opt:pregen=0 2>&1 | tail -3
WORKLOAD=c4 timeout 1500 bash tests/ab_variants.sh opt:pregen=0 2>&1 | tail -3
export MESHGEN_LIB=$PWD/reinforcementlearning4meshgeneration_b200/lib/variants/trace.so
timeout 300 python tests/trace_items.py c3 gpurun_out/trace_c3 6 > gpurun_out/trace_c3.log 2>&1; tail -3 gpurun_out/trace_c3.log
unset MESHGEN_LIB
