cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'attn_bwd|ln_res_drop' -c 8 -o gpurun_out/r02c_full -f python tools/prof_step.py --layers 2 --serial-wgrad > gpurun_out/r02c_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/r02c*
