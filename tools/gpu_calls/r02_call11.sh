cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_decode_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02k_pytest.log | cut -c1-300
timeout 400 python tools/decode_bench.py > gpurun_out/r02k_decode.log 2>&1; echo "decode rc=$?"; tail -8 gpurun_out/r02k_decode.log | cut -c1-250
