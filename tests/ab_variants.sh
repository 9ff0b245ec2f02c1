# A/B of tuning variants in one gpurun call (profiling aid): build variant libraries into
# reinforcementlearning4meshgeneration_b200/lib/variants/<name>.so (nvcc ... -DMG_...=...), then
#   gpurun -- 'bash tests/ab_variants.sh <name> <name> ...'
for v in base "$@"; do
  if [ $v = base ]; then unset MESHGEN_LIB; else export MESHGEN_LIB=$PWD/reinforcementlearning4meshgeneration_b200/lib/variants/$v.so; fi
  python bench.py --no-cpu-baseline --phase-times --steps 500 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$v.json"))
    print("$v", "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], d.get("phase_times"), "e2e %.4g" % d["e2e"]["value"])
except Exception as ex:
    print("$v", "failed", ex)
PY
done
