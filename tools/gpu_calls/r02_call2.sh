# validate the ring LN kernels and the persistent attention backward; bench
cd /root/repo; mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02b_pytest.log | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02b_bench.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/r02b_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print(json.dumps(d['roofline']['families_ms_per_step']))"
timeout 300 python tools/attn_bench.py > gpurun_out/r02b_attn_bench.log 2>&1; tail -8 gpurun_out/r02b_attn_bench.log
