for v in base ns20 ns24 ns28 ns32; do
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
