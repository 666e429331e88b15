timeout 900 python -m pytest tests/test_gpu_qgemm.py -x -q 2>&1 | tail -3
C=12500000,96,u8,cosine,10,1024,gemm,3
for D in 0 16 1; do echo "debug $D"; EVDB_QGEMM_DEBUG=$D python tools/sweep.py $C 2>&1 | tail -1; done
python tools/sweep.py 12500000,96,u8,cosine,10,8,gemm,3 12500000,96,u8,cosine,10,64,gemm,3 12500000,96,u8,cosine,100,1024,gemm,3 2>&1 | tail -3
python tools/sweep.py 1000000,768,u8,cosine,10,1024,gemm,3 4000000,128,u8,cosine,10,1024,gemm,3 2>&1 | tail -2
