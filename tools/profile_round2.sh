# Round-2 captures (one B200, through gpurun): bench first, then launch lists, then ncu --set full of the top kernels.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err || exit 1
for m in fp32 tf32 simt; do
  python bench.py --math $m --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/r2_plain_$m.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2_$m.csv python bench.py --math $m --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/r2_ncu_l_$m.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:umma_conv_gather_f16 -c 10 -o gpurun_out/prof_f16_gather_r2 -f python bench.py --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/r2_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_conv_wgrad_f16 -c 5 -o gpurun_out/prof_f16_wgrad_r2 -f python bench.py --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/r2_ncu_f2.log 2>&1
ncu --set full --clock-control none -k "regex:stage_f16|bn_reduce|bn_apply|bn_bwd_dx|pool_fwd|pool_bwd|multi_update|filter_stage" -c 36 -o gpurun_out/prof_membound_r2 -f python bench.py --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/r2_ncu_f3.log 2>&1
ncu --set full --clock-control none -k regex:smallc -c 2 -o gpurun_out/prof_smallc_r2 -f python bench.py --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/r2_ncu_f4.log 2>&1
# gpurun copies back at most 64 MiB: summarise on the box, keep only the two tensor-core captures
for k in f16_gather f16_wgrad membound smallc; do python tools/ncu_summary.py gpurun_out/prof_${k}_r2.ncu-rep > gpurun_out/r2_ncu_$k.txt 2>&1; done
rm -f gpurun_out/prof_membound_r2.ncu-rep gpurun_out/prof_smallc_r2.ncu-rep
ls -la gpurun_out/ | head -40; du -sh gpurun_out
