# Round-2 evidence call: GPU tests, default bench line, launch list of one real step, ncu --set full of the hot kernels (2-layer step).
cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_1gpu.err | cut -c1-300
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py --serial-wgrad > gpurun_out/r02_launches.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_|ln_res' -o gpurun_out/r02_full -f python tools/prof_step.py --layers 2 --serial-wgrad > gpurun_out/r02_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | head -30
