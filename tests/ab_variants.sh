# A/B of tuning variants in one gpurun call (profiling aid).  Variant libraries are built into
# reinforcementlearning4meshgeneration_b200/lib/variants/<name>.so (nvcc ... -DMG_...=...); a variant name of the form
# opt:<k=v,...> runs the default library with MESHGEN_OPTIONS=<k=v,...>.
#   gpurun -- 'bash tests/ab_variants.sh <name> <name> ...'
WL=${WORKLOAD:-c3}
for v in base "$@"; do
  unset MESHGEN_LIB MESHGEN_OPTIONS
  case $v in
    base) ;;
    opt:*) export MESHGEN_OPTIONS=${v#opt:} ;;
    *@*) export MESHGEN_LIB=$PWD/reinforcementlearning4meshgeneration_b200/lib/variants/${v%%@*}.so; export MESHGEN_OPTIONS=${v#*@} ;;
    *) export MESHGEN_LIB=$PWD/reinforcementlearning4meshgeneration_b200/lib/variants/$v.so ;;
  esac
  tag=$(echo $v | tr ':=,@' '____')
  python bench.py --workload $WL --no-cpu-baseline --steps 400 --warmup 10 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$tag.json"))
    print("$v", "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], [round(k["ms"] * 1e3, 1) for k in d["roofline"]["per_kernel"]], "e2e %.4g" % d["e2e"]["value"])
except Exception as ex:
    print("$v", "failed", ex)
PY
done
