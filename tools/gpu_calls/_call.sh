cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_parity_gpu.py::test_full_size_properties_bert_base_b64 > gpurun_out/r02_c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c19_pytest.log | cut -c1-300
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c19_bench.json 2> gpurun_out/r02_c19_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_c19_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r02_c19_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'], d.get('optimizer',{}).get('ms_per_step'));print(json.dumps(d['roofline']['families_ms_per_step']))"
python tools/gemm_vs_cublas.py gpurun_out/r02_gemm_vs_cublas.md > gpurun_out/r02_c19_gemm.log 2>&1; echo "gemm rc=$?"
