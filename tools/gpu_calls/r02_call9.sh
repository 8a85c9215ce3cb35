# final 1-GPU evidence of round 2: GPU tests, smoke, bench lines (caption, vqa, reference arm), decode and optimizer micro-benchmarks
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02i_pytest.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02i_smoke.log
timeout 600 python bench.py > gpurun_out/r02i_bench_1gpu.json 2> gpurun_out/r02i_bench_1gpu.err; echo "bench rc=$?"
timeout 400 python bench.py --config vqa --no-cpu-baseline > gpurun_out/r02i_vqa_1gpu.json 2> gpurun_out/r02i_vqa_1gpu.err; echo "vqa rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02i_ref.json 2> gpurun_out/r02i_ref.err; echo "ref rc=$?"
for f in bench_1gpu vqa_1gpu; do python -c "
import json;d=json.load(open('gpurun_out/r02i_$f.json'));print('$f',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'eager',d.get('eager',{}).get('value'),'opt',d.get('optimizer',{}).get('ms_per_step'),d['clocks']['sm_mhz'],d['clocks']['reasons'])"; done
timeout 300 python tools/decode_bench.py > gpurun_out/r02i_decode.log 2>&1; echo "decode rc=$?"; tail -8 gpurun_out/r02i_decode.log | cut -c1-200
timeout 200 python tools/opt_bench.py > gpurun_out/r02i_opt.log 2>&1; echo "opt rc=$?"; tail -4 gpurun_out/r02i_opt.log | cut -c1-200
