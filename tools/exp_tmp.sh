timeout 900 python -m pytest tests/test_gpu_qgemm.py -x -q 2>&1 | tail -3
for B in 8 16 64 128 200; do python tools/sweep.py 12500000,96,u8,cosine,10,$B,gemm,5 2>&1 | tail -1; done
for B in 4 8 64; do python tools/sweep.py 1000000,768,u8,cosine,10,$B,gemm,5 2>&1 | tail -1; done
