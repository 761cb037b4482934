timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r02v2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v2_pytest.log; tail -3 gpurun_out/r02v2_pytest.log
