set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err || exit 1
for m in fp32 tf32 simt; do
  python bench.py --math $m --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/plain_$m.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1_$m.csv python bench.py --math $m --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/ncu_l_$m.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:umma_conv_gather -c 5 -o gpurun_out/prof_x3_gather_r1 -f python bench.py --math fp32 --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_conv_wgrad -c 5 -o gpurun_out/prof_x3_wgrad_r1 -f python bench.py --math fp32 --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_conv_gather -c 5 -o gpurun_out/prof_umma_gather_r1 -f python bench.py --math tf32 --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_conv_wgrad -c 5 -o gpurun_out/prof_umma_wgrad_r1 -f python bench.py --math tf32 --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:smallc -c 2 -o gpurun_out/prof_smallc_r1 -f python bench.py --math fp32 --no-tf32 --no-extras --steps 2 --warmup 3 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/*.ncu-rep
