timeout 900 python -m pytest tests/test_gpu_qgemm.py -x -q 2>&1 | tail -2
for c in 12500000,96,u8,cosine,10,1024 12500000,96,u8,cosine,100,1024 12500000,96,u8,cosine,10,128 12500000,96,u8,cosine,10,64 4000000,256,u8,cosine,10,1024 1000000,768,u8,cosine,10,1024 2000000,1536,u8,cosine,10,1024; do python tools/sweep.py $c,gemm,5 2>&1 | tail -1 | cut -c1-140; done
