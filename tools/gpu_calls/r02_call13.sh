# 4 GPUs, configs[1] shape, default protocol (30 steps): a weak-scaling point between the 2- and 8-GPU runs
cd /root/repo; mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r02m_caption_4gpu.json 2> gpurun_out/r02m_caption_4gpu.err; echo "cap4 rc=$?"
python -c "
import json;txt=[l for l in open('gpurun_out/r02m_caption_4gpu.json') if l.startswith('{')][0];d=json.loads(txt);print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d.get('comm'),d['clocks'])"
