# 1 GPU, configs[1], 2000 timed steps: the sustained single-GPU number (power / clocks) that the 8-GPU sustained run should be compared with
cd /root/repo; mkdir -p gpurun_out
timeout 300 python bench.py --steps 2000 --warmup 5 --no-cpu-baseline > gpurun_out/r02l_caption_1gpu_sustained.json 2> gpurun_out/r02l.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02l_caption_1gpu_sustained.json'));print(d['value'],d['ms_per_step'],d['step_ms'],'e2e',d['e2e']['value'],d['clocks'])"
