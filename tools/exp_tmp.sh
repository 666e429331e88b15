for B in 4 6 8; do python tools/sweep.py 12500000,96,u8,cosine,10,$B,gemm,5 12500000,96,u8,cosine,10,$B,scan,5 2>&1 | tail -2 | cut -c1-120; done
for B in 3 4 6; do python tools/sweep.py 1000000,768,u8,cosine,10,$B,gemm,5 1000000,768,u8,cosine,10,$B,scan,5 2>&1 | tail -2| cut -c1-120; done
for B in 4 6; do python tools/sweep.py 4000000,256,u8,cosine,10,$B,gemm,5 4000000,256,u8,cosine,10,$B,scan,5 2>&1 | tail -2| cut -c1-120; done
python tools/sweep.py 12500000,96,u8,cosine,100,1024,gemm,5 2000000,1536,u8,cosine,10,1024,gemm,5 12500000,96,u8,cosine,10,128,gemm,5 12500000,96,u8,cosine,10,16,gemm,5 1000000,768,u8,cosine,10,64,gemm,5  2>&1 | tail -5 | cut -c1-150
timeout 1200 python -m pytest tests/test_gpu_qgemm.py tests/test_gpu_fullsize.py tests/test_gpu_mstore.py -x -q -k "i8 or config4 or quantized" 2>&1 | tail -3
