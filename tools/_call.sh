cd /root/repo
timeout 600 python -m pytest tests/test_table_grads_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c15_pytest.log | cut -c1-300
run() { name=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/r02_c15_$name.json 2> gpurun_out/r02_c15_$name.err; python -c "
import json
l=[x for x in open('gpurun_out/r02_c15_$name.json') if x.startswith('{')]
d=json.loads(l[-1]);print('N2 $name',d['value'],d['ms_per_step'])"; }
run base VLP_X=0
run ctas8_res8 NCCL_MAX_CTAS=8 VLP_DP_RESERVED_SMS=8
run ctas4_res4 NCCL_MAX_CTAS=4 VLP_DP_RESERVED_SMS=4
run ctas16_res16 NCCL_MAX_CTAS=16 VLP_DP_RESERVED_SMS=16
run ctas2_res2 NCCL_MAX_CTAS=2 VLP_DP_RESERVED_SMS=2
run groups_1_1_2_4_4 VLP_DP_GROUPS=1,1,2,4,4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --config vqa > gpurun_out/r02_c15_n2_vqa.json 2> gpurun_out/r02_c15_n2_vqa.err; echo "n2 vqa rc=$?"; python -c "
import json,sys
l=[x for x in open('gpurun_out/r02_c15_n2_vqa.json') if x.startswith('{')]
d=json.loads(l[-1]);print('N2 vqa',d['value'],d['ms_per_step'],d['e2e']['value'], d.get('optimizer',{}).get('ms_per_step'), d.get('comm'))"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config vqa > gpurun_out/r02_c15_n1_vqa.json 2> gpurun_out/r02_c15_n1_vqa.err; python -c "
import json;d=json.load(open('gpurun_out/r02_c15_n1_vqa.json'));print('N1 vqa',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('optimizer',{}).get('ms_per_step'))"
