#!/bin/bash
# A/B of run-time knobs on the headline workload (device-resident arm): draw-stage chunk size, wave pipelines, quantum.
B="python bench.py --skip-e2e --no-cpu-baseline --complete-photons 0 --steps 10 --warmup 3"
run() { echo -n "$1: "; shift; env "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g'%d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['per_kernel'].items()})"; }
run default $B
run chunk512 LART_GPU_DRAW_CHUNK=512 $B
run chunk2048 LART_GPU_DRAW_CHUNK=2048 $B
run chunk4096 LART_GPU_DRAW_CHUNK=4096 $B
run streams4 $B --streams 4
run streams8 $B --streams 8
run streams12 $B --streams 12
run quantum64 $B --quantum 64
