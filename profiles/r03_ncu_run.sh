# Round 3 captures (one B200, gpurun): the launch list of the bench command and `ncu --set full` of the CTA-pair FaceNeRF kernel, of the
# fused small kernels of inerf_render_rays_fused and of the training kernels.  Every program runs plain first (must exit 0 without ncu).
set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
python profiles/r02_kernels.py render > gpurun_out/r03_k_render.log 2>&1 || exit 1
python profiles/r02_kernels.py train > gpurun_out/r03_k_train.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-extra --video-frames 0 > gpurun_out/r03_plain_bench.json 2> gpurun_out/r03_plain_bench.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r03_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-extra --video-frames 0 > gpurun_out/r03_ncu_bench.log 2>&1
$NCU --set full --import-source on -k regex:mlp_bf16_kernel -s 2 -c 2 -o /tmp/r03_render python profiles/r02_kernels.py render > gpurun_out/r03_ncu_render.log 2>&1
$NCU --set full --import-source on -k regex:"render_setup|composite_sample|composite_final" -s 3 -c 3 -o /tmp/r03_small python profiles/r02_kernels.py render > gpurun_out/r03_ncu_small.log 2>&1
$NCU --set full -k regex:"mlp_bf16_kernel|bwd_chain|dw_kernel" -s 18 -c 6 -o /tmp/r03_train python profiles/r02_kernels.py train > gpurun_out/r03_ncu_train.log 2>&1
for f in render small train; do
  ncu -i /tmp/r03_$f.ncu-rep --page raw --csv > gpurun_out/r03_${f}_raw.csv 2>/dev/null
done
ncu -i /tmp/r03_render.ncu-rep --page source --csv > gpurun_out/r03_render_source.csv 2>/dev/null
ls -la gpurun_out/ | grep r03_
