# configs[3] (VQA, batch 128 per GPU) at 1 and 2 GPUs, configs[1] at 2 GPUs
cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/r02d_smi.txt
timeout 400 python bench.py --config vqa --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_vqa_1gpu.json 2> gpurun_out/r02d_vqa_1gpu.err; echo "vqa1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config vqa --steps 50 --warmup 5 > gpurun_out/r02d_vqa_2gpu.json 2> gpurun_out/r02d_vqa_2gpu.err; echo "vqa2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02d_caption_2gpu.json 2> gpurun_out/r02d_caption_2gpu.err; echo "cap2 rc=$?"
for f in vqa_1gpu vqa_2gpu caption_2gpu; do python -c "
import json,sys;d=json.load(open('gpurun_out/r02d_$f.json'));print('$f',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('comm'),d['step_ms'])"; done
tail -2 gpurun_out/r02d_*.err | cut -c1-300
