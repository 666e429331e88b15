for CO in 0 1; do echo "COARSE=$CO"; for c in 12500000,96,u8,cosine,10,1024 12500000,96,u8,cosine,100,1024 12500000,96,u8,cosine,10,64 12500000,96,u8,cosine,10,8 4000000,128,u8,cosine,10,1024; do EVDB_QGEMM_COARSE=$CO python tools/sweep.py $c,gemm,5 2>&1 | tail -1 | cut -c1-140; done; done
timeout 900 python -m pytest tests/test_gpu_qgemm.py -x -q 2>&1 | tail -2
