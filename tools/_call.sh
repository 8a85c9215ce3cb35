cd /root/repo
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c13_n1.json 2> gpurun_out/r02_c13_n1.err; echo "n1 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_c13_n1.json'));print('N1',d['value'],d['ms_per_step'],d['e2e']['value'], d.get('optimizer',{}).get('ms_per_step'))"
for cfg in caption vqa; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --config $cfg > gpurun_out/r02_c13_n2_$cfg.json 2> gpurun_out/r02_c13_n2_$cfg.err; echo "n2 $cfg rc=$?"; tail -2 gpurun_out/r02_c13_n2_$cfg.err | cut -c1-300; python -c "
import json,sys
l=[x for x in open('gpurun_out/r02_c13_n2_$cfg.json') if x.startswith('{')]
d=json.loads(l[-1]);print('N2 $cfg',d['value'],d['ms_per_step'],d['e2e']['value'], d.get('optimizer',{}).get('ms_per_step'), d.get('comm'))"
done
VLP_DP_RESERVED_SMS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/r02_c13_n2_res8.json 2> gpurun_out/r02_c13_n2_res8.err; python -c "
import json
l=[x for x in open('gpurun_out/r02_c13_n2_res8.json') if x.startswith('{')]
d=json.loads(l[-1]);print('N2 reserve8',d['value'],d['ms_per_step'])"
VLP_DP_GROUPS=3,3,3,3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/r02_c13_n2_g3333.json 2> gpurun_out/r02_c13_n2_g3333.err; python -c "
import json
l=[x for x in open('gpurun_out/r02_c13_n2_g3333.json') if x.startswith('{')]
d=json.loads(l[-1]);print('N2 groups 3,3,3,3',d['value'],d['ms_per_step'])"
