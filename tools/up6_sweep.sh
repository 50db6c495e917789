#!/bin/bash
# up6_tc_kernel parameter sweep (ring depths, L2 prefetch distance, pipeline-stage skips): bench.py's per-stage CUDA-event spans.
# usage: gpurun -- 'bash tools/up6_sweep.sh > gpurun_out/up6_sweep.txt'
# configurations: the arguments (one quoted "VAR=.. VAR=.." string each), else the stage-skip sweep and the two epilogue forms
if [ $# -gt 0 ]; then CFGS=("$@"); else CFGS=("SRT_UP6_DBG=0" "SRT_UP6_DBG=1" "SRT_UP6_DBG=2" "SRT_UP6_DBG=4" "SRT_UP6_DBG=8" "SRT_UP6_DBG=15" "SRT_UP6_PAIR=0" "SRT_UP6_PAIR=1" "SRT_UP6_LO8=0"); fi
run() {
  env SRT_BENCH_STAGE_SKIPS=1 "$@" timeout 200 python bench.py --no-extras --steps 10 --warmup 3 2>>gpurun_out/up6_sweep.err | grep '^{' | python -c '
import json,sys
d=json.loads(sys.stdin.read()); k=[x for x in d["roofline_all"] if x["stage"]=="up6+up7"][0]
print("%.4f ms up6+up7  %.3f ms/step  err %s" % (k["ms"], d["ms_per_step"], ["%.1e"%e for e in d["parity"]["stem_rms_err"]]))'
}
for cfg in "${CFGS[@]}"; do
  echo -n "$cfg : "; run $cfg
done
