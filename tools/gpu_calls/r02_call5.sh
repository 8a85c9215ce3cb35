cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02e_pytest_graph.log 2>&1; echo "pytest graph rc=$?"; tail -15 gpurun_out/r02e_pytest_graph.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02e_bench.err | cut -c1-400
python -c "
import json;d=json.load(open('gpurun_out/r02e_bench.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'eager',d.get('eager'),'opt',d.get('optimizer',{}).get('ms_per_step'), d['gpu_launches']);print(json.dumps(d['roofline']['families_ms_per_step'])); print(d['roofline']['hbm_kernels'])"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02e_pytest.log | cut -c1-300
