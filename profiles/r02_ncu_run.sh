set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
# plain runs first (must exit 0 without ncu)
python profiles/r02_kernels.py render > gpurun_out/r02_k_render.log 2>&1 || exit 1
python profiles/r02_kernels.py f16x2 > gpurun_out/r02_k_f16x2.log 2>&1 || exit 1
python profiles/r02_kernels.py sampling > gpurun_out/r02_k_sampling.log 2>&1 || exit 1
python profiles/r02_kernels.py train > gpurun_out/r02_k_train.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-extra --video-frames 0 > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || exit 1
# launch list of the bench command
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-extra --video-frames 0 > gpurun_out/r02_ncu_bench.log 2>&1
# full captures
$NCU --set full --import-source on -k regex:mlp_bf16_kernel -s 2 -c 2 -o /tmp/r02_render python profiles/r02_kernels.py render > gpurun_out/r02_ncu_render.log 2>&1
$NCU --set full --import-source on -k regex:mlp_f16x2_kernel -s 2 -c 2 -o /tmp/r02_f16x2 python profiles/r02_kernels.py f16x2 > gpurun_out/r02_ncu_f16x2.log 2>&1
$NCU --set full --import-source on -k regex:"importance_rng_kernel|sample_coarse_kernel|composite" -s 4 -c 6 -o /tmp/r02_small python profiles/r02_kernels.py render > gpurun_out/r02_ncu_small.log 2>&1
$NCU --set full -k regex:"mlp_bf16_kernel|bwd_chain|dw_kernel" -s 18 -c 6 -o /tmp/r02_train python profiles/r02_kernels.py train > gpurun_out/r02_ncu_train.log 2>&1
for f in render f16x2 small train; do
  ncu -i /tmp/r02_$f.ncu-rep --page raw --csv > gpurun_out/r02_${f}_raw.csv 2>/dev/null
done
ncu -i /tmp/r02_f16x2.ncu-rep --page source --csv > gpurun_out/r02_f16x2_source.csv 2>/dev/null
ncu -i /tmp/r02_small.ncu-rep --page source --csv > gpurun_out/r02_small_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -20
