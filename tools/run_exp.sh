# usage: bash tools/run_exp.sh <tag> [variants...]   (variants: "NAME:ENV=VAL ENV2=VAL2")
tag=$1; shift
B="timeout 250 python bench.py --steps 10 --warmup 3 --cpu-sample 200"
run() {  # name, env...
  name=$1; shift
  env "$@" $B > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_${tag}_$name.json").read().strip().splitlines()[-1])
    print("${tag}_$name", "qps %.4g ms/step %.3f kernel_ms %.4f frac %.3f e2e %.4g recall %.4f"%(l["value"],l["ms_per_step"],l["roofline"]["kernel_ms"],l["roofline"]["frac"],l["e2e"]["value"],l["recall_at_10"]))
except Exception as e:
    print("${tag}_$name FAILED", e)
PY
  grep "tc batch" gpurun_out/bench_${tag}_$name.err | tail -1
}
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  run $name $envs
done
