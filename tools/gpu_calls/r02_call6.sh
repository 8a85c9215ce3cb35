# 16-warp GEMM epilogue validation (tests, 1-GPU bench, per-shape table) + CUDA-graph capture of the data-parallel step at 2 GPUs
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02f_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/r02f_bench.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'eager',d.get('eager',{}).get('value'));print(json.dumps(d['roofline']['families_ms_per_step']))"
timeout 300 python tools/gemm_vs_cublas.py gpurun_out/r02f_gemm_vs_cublas.md > gpurun_out/r02f_gemm.log 2>&1; echo "gemm rc=$?"; tail -16 gpurun_out/r02f_gemm_vs_cublas.md | cut -c1-200
VLP_BENCH_GRAPH_DP=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02f_caption_2gpu_graph.json 2> gpurun_out/r02f_caption_2gpu_graph.err; echo "cap2 graph rc=$?"
VLP_BENCH_GRAPH_DP=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config vqa --steps 50 --warmup 5 > gpurun_out/r02f_vqa_2gpu_graph.json 2> gpurun_out/r02f_vqa_2gpu_graph.err; echo "vqa2 graph rc=$?"
for f in caption_2gpu_graph vqa_2gpu_graph; do python -c "
import json;txt=[l for l in open('gpurun_out/r02f_$f.json') if l.startswith('{')][0];d=json.loads(txt);print('$f',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'eager',d.get('eager',{}).get('value'),d.get('comm'))"; tail -4 gpurun_out/r02f_$f.err | cut -c1-300; done
