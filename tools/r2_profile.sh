#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): bench lines of every workload, the ncu launch list and the
# `--set full` capture of the wave kernels.  Everything lands in gpurun_out/ (scratch); the summaries are copied to profiles/.
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
for w in sphere_peel_tau1e7_coreskip vel_effect_peel vel_effect_peel_as_shipped slab_tau1e7 sphere_peel_tau1e4 sphere_octant_tau1e7 sphere_quadrant_tau1e7 box_periodic_tau1e7 clump_sphere_fcov5 amr_sphere_tau1e4 amr_sphere_tau1e7 clump_overlap_fcov3 sphere_tau1e7_calcJP_radial sphere_tau1e7_calcJP_cells box_shear_tau1e7 plane_atmosphere_tau1e6 spherical_atmosphere_tau1e6; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 5 --complete-photons 0 > $O/r2_bench_$w.json 2> $O/r2_bench_$w.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --flags 4 --skip-e2e --no-cpu-baseline --complete-photons 0 > $O/r2_bench_mono.json 2>/dev/null
timeout 200 python bench.py --steps 10 --warmup 3 --flags 16 --skip-e2e --no-cpu-baseline --complete-photons 0 > $O/r2_bench_serial_rejection.json 2>/dev/null
timeout 200 python bench.py --steps 10 --warmup 3 --flags 128 --skip-e2e --no-cpu-baseline --complete-photons 0 > $O/r2_bench_speculative_rejection.json 2>/dev/null
timeout 200 python bench.py --steps 10 --warmup 3 --flags 2 --skip-e2e --no-cpu-baseline --complete-photons 0 > $O/r2_bench_no_warp_agg.json 2>/dev/null
# ncu only after the same command has run clean above
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --complete-photons 0 --streams 1 > $O/r2_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_" --launch-skip 420 --launch-count 6 -f -o $O/r2_full python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --complete-photons 0 --streams 1 > $O/r2_ncu_full.log 2>&1
ls -la $O | tail -30
