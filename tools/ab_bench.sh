#!/bin/bash
# A/B of prebuilt libraries on ONE box (boxes differ by ~1%, runs on one box by ~0.01%):
#   tools/ab_bench.sh ab/lib_a.so ab/lib_b.so ab/lib_b.so@SOME_ENV_KNOB ...   (two alternating rounds, C2 device-resident)
for round in 1 2; do
  for spec in "$@"; do
    lib=${spec%@*}; knob=""; [ "$spec" != "$lib" ] && knob=${spec#*@}
    cp "$lib" qec_ldpc_b200/lib/libqldpc_b200.so
    env ${knob:+$knob=1} python bench.py --config C2 --steps 3 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$spec', 'frac %.4f' % d['roofline']['frac'], 'value %.4g' % d['value'], 'bp_x %.2f bp_z %.2f' % (k['bp_x'], k['bp_z']))"
  done
done
