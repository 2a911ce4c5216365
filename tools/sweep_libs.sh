#!/bin/sh
# experiments: run the ME parity tests, a short bench and the content sweep for every variant library build/lib_*.so
for lib in "" build/lib_*.so; do
  [ -n "$lib" ] && export P64B_LIB=$PWD/$lib
  echo "== ${lib:-product}"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "me_matches or adversarial_content_matches" 2>&1 | tail -1
  python bench.py --steps 30 --warmup 4 --no-rate-control --no-sustained 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline_kernels']['me_search_kernel']
print('value %.0f e2e %.0f me_ms %.4f mb_ms %.4f exec_share %.3f exec_frac %.3f' % (d['value'], d['e2e']['value'], r['avg_launch_ms'], d['roofline_kernels']['mb_encode_kernel']['avg_launch_ms'], r['executed']['share_of_algorithmic'], r['executed']['frac_of_peak']))"
  python tools/me_content_sweep.py 2>/dev/null | cut -c1-140
done
