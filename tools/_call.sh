cd /root/repo
python tools/tune_tiles.py > gpurun_out/r02_c3_tiles.log 2>&1; echo "tiles rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -o gpurun_out/r02_gemm_prof -f python tools/gemm_probe.py > gpurun_out/r02_c3_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
