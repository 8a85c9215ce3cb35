# 8 GPUs, CC-pretrain shape, 2000 steps: A/B of the tied-embedding gradient path (early decoder all-reduce + all-gathered lookup rows instead of the dense tail all-reduce)
cd /root/repo; mkdir -p gpurun_out
s=$(date +%s)
VLP_DP_SPARSE_EMB=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --config ccmix --steps 2000 --warmup 5 > gpurun_out/r02j_ccmix_8gpu_sparse.json 2> gpurun_out/r02j_ccmix_8gpu_sparse.err; echo "ccmix8 sparse rc=$? after $(( $(date +%s) - s )) s"
python -c "
import json;txt=[l for l in open('gpurun_out/r02j_ccmix_8gpu_sparse.json') if l.startswith('{')][0];d=json.loads(txt);print(d['value'],d['ms_per_step'],d['step_ms'],'e2e',d['e2e']['value'],d.get('comm'),d['clocks'])"; tail -3 gpurun_out/r02j_ccmix_8gpu_sparse.err | cut -c1-300
