#!/bin/bash
# Builds imdct_sparse_kernel variants ON THE GPU BOX and times the decode of the hour-long bench signal.
# usage (under gpurun): bash tools/imdct_sweep.sh "<flags variant 1>" "<flags variant 2>" ...
# flags: -DGLC_IMDCT_BN=256|512 -DGLC_IMDCT_RW=1|2 -DGLC_IMDCT_KC=16|32 -DGLC_IMDCT_RING=2..6 -DGLC_IMDCT_MINB=1..3
#        -DGLC_IMDCT_ROWMASK=1 -DGLC_IMDCT_X2=0|1 (packed multiply-adds, default 1; the shapes other than 2 rows per warp and
#        ROWMASK need -DGLC_IMDCT_X2=0) -DGLC_MDCT_X2=0|1
cd "$(dirname "$0")/../gapless_lossy_codec_b200/csrc" || exit 1
unset CC CXX
for v in "$@"; do
  rm -f build/glc_exact_gemm.o build/glc_codec_kernels.o build/glc_api.o build/glc_tables.o
  make -j16 -s EXTRA="$v" > /tmp/mk.log 2>&1 || { echo "BUILD FAILED: $v"; tail -5 /tmp/mk.log; continue; }
  (cd ../.. && timeout 300 python bench.py --no-cpu-baseline --no-flac --no-fast-side > /tmp/b.json 2> /tmp/b.err && python -c "
import json,sys;d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1]);k=d['roofline']['kernel_ms_per_step'];print('VARIANT',sys.argv[1],'| imdct',round(k['imdct_exact'],3),'dequant',round(k['dequant'],3),'step',round(d['ms_per_step'],2))" "$v") || { echo "RUN FAILED: $v"; tail -3 /tmp/b.err; }
done
