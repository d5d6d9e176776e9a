#!/bin/bash
# Runs bench.py for the given configurations on N GPUs of this box and keeps the JSON lines under gpurun_out/.
#   tools/run_configs.sh N C2 C3 C5 ...        (N = 1: plain python; N > 1: torch.distributed.run, one rank per GPU)
N=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  out=gpurun_out/bench_${cfg}_${N}gpu.json
  if [ "$N" = "1" ]; then
    python bench.py --config $cfg --steps 3 --warmup 3 > $out.raw 2> $out.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus $N --config $cfg --steps 3 --warmup 3 > $out.raw 2> $out.err
  fi
  echo "$cfg x$N rc=$?"
  grep '^{' $out.raw > $out; rm -f $out.raw
  python - <<PY
import json
try:
    d = json.load(open("$out"))
    e = d["e2e"]
    print("  value %.4g  roofline %s %.3f  e2e %.4g (%.2f of value, host roofline %.2f)  u8 %.4g  devgen %.4g" % (
        d["value"], d["roofline"]["bound"], d["roofline"]["frac"], e["value"], e["vs_device_value"],
        e["host_mem_roofline"]["frac"], e["u8_patterns"]["value"], e["device_generated"]["value"]))
except Exception as ex:
    print("  no line:", ex)
PY
  tail -c 300 $out.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -3
done
