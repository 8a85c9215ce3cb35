# 2 GPUs: the data-parallel step under CUDA-graph replay must print its line AND exit (round-2 hang at destroy_process_group)
cd /root/repo; mkdir -p gpurun_out
s=$(date +%s)
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02g_caption_2gpu.json 2> gpurun_out/r02g_caption_2gpu.err; echo "cap2 graph rc=$? after $(( $(date +%s) - s )) s"
python -c "
import json;txt=[l for l in open('gpurun_out/r02g_caption_2gpu.json') if l.startswith('{')][0];d=json.loads(txt);print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'eager',d.get('eager',{}).get('value'),d.get('comm'))"; tail -3 gpurun_out/r02g_caption_2gpu.err | cut -c1-300
s=$(date +%s)
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02g_ref_2gpu.json 2> gpurun_out/r02g_ref_2gpu.err; echo "ref arm rc=$? after $(( $(date +%s) - s )) s"; tail -c 400 gpurun_out/r02g_ref_2gpu.json
