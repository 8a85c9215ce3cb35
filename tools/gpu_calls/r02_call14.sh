cd /root/repo; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_graph_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02n_pytest.log | cut -c1-300
