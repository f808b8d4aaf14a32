#!/bin/bash
# two depth maps in flight: repeated runs of the bench (a hang shows up as a trap in a bounded mbarrier wait) + core dump on failure
export CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_COREDUMP_FILE=/tmp/gpucore CUDA_COREDUMP_GENERATION_FLAGS="skip_global_memory,skip_local_memory,skip_constbank_memory"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py -q -x 2>&1 | tail -2
for i in 1 2 3 4 5 6; do
  timeout 300 python bench.py --steps 20 --warmup 3 --streams 2 --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err
  if [ -s /tmp/b.json ]; then python -c "import json; b=json.load(open('/tmp/b.json')); print('ok', b['value'], b['e2e']['value'])"; else echo "FAILED run $i"; break; fi
done
timeout 300 python bench.py --steps 10 --warmup 3 --streams 1 --no-cpu-baseline > /tmp/b1.json 2>/dev/null; python -c "import json; b=json.load(open('/tmp/b1.json')); print('1 stream', b['value'], b['e2e']['value'])"
timeout 300 python bench.py --workload tt --steps 8 --warmup 3 --no-cpu-baseline > /tmp/b2.json 2>/dev/null; python -c "import json; b=json.load(open('/tmp/b2.json')); print('tt 2 streams', b['value'], b['e2e']['value'])"
if ls /tmp/gpucore* >/dev/null 2>&1; then
  f=$(ls /tmp/gpucore* | head -1)
  printf 'set pagination off\ninfo cuda kernels\ninfo cuda warps\n' > /tmp/gdbcmds
  timeout 300 cuda-gdb -batch -ex "target cudacore $f" -x /tmp/gdbcmds 2>&1 | grep -v "^\[New\|^warning" | head -60
fi
