# configs[4]: CC-pretrain shape (mixed s2s / bi masks), 8 x B200, sustained-throughput protocol: 2000 timed steps, per-step percentiles, clock / power trace
cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r02h_clocks.csv &
SMI=$!
s=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config ccmix --steps 2000 --warmup 5 > gpurun_out/r02h_ccmix_8gpu.json 2> gpurun_out/r02h_ccmix_8gpu.err; echo "ccmix8 rc=$? after $(( $(date +%s) - s )) s"
kill $SMI
python -c "
import json;txt=[l for l in open('gpurun_out/r02h_ccmix_8gpu.json') if l.startswith('{')][0];d=json.loads(txt);print(d['value'],d['ms_per_step'],d['step_ms'],'e2e',d['e2e']['value'],'eager',d.get('eager',{}).get('value'),d.get('comm'),d['clocks'])"; tail -3 gpurun_out/r02h_ccmix_8gpu.err | cut -c1-300; wc -l gpurun_out/r02h_clocks.csv
