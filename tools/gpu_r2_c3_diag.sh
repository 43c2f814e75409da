mkdir -p gpurun_out/r2
timeout -s KILL 600 python tools/diag_c3.py > gpurun_out/r2/diag_c3_v1.txt 2>&1
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider --maxfail=8 > gpurun_out/r2/test_all_c3v1.txt 2>&1
tail -12 gpurun_out/r2/test_all_c3v1.txt
