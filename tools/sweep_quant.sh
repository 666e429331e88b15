# Same-box A/B of the quantized scans.  Variants are selected by environment (EVDB_SCAN_TMA=0/1, ...)
# or by a second build: python tools/build_variant.py build_ab/libevdb_q3.so -DEVDB_QPLANES=3
SH="12500000,96,u8,cosine,10,1 12500000,128,u8,cosine,10,1 1000000,1536,u4,cosine,10,1 1000000,768,u8,cosine,10,1 2000000,1536,u8,cosine,10,1 8000000,64,u4,cosine,10,1 4000000,256,u8,cosine,10,1 300000,6144,u8,cosine,10,1 12500000,96,u8,cosine,10,8 12500000,96,u8,cosine,100,1"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_staged or scaled_configs or quant or u8 or u4 or codes" 2>&1 | tail -2
for cfg in "EVDB_LIB_PATH=build_ab/libevdb_q3.so" "EVDB_SCAN_DEBUG=1"; do
  echo "== $cfg"
  env $cfg EVDB_SCAN_DEBUG=1 timeout 300 python tools/sweep.py $SH 2>&1 | python -c "
import sys, json
plan=''
for l in sys.stdin:
    if l.startswith('[evdb scan]'): plan=l.split('plan:')[1].strip(); continue
    if l.startswith('{'):
        r=json.loads(l); print(r['case'][:22].ljust(24), r['kernel_ms'], r.get('gbs'), plan); plan=''
"
done
