timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_kernels.py tests/test_gpu_next_rows.py -q -x 2>&1 | grep -E "^E   |passed|failed|FAILED|rror" | cut -c1-250 | head -5
timeout 100 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['bwd']; print('step %.3f  dgrad %.3f  wgrad %.3f fwd %.3f  value %.0f clocks %s launches %d' % (d['ms_per_step'], r['data_kernel_ms'], r['weight_kernel_ms'], d['roofline']['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['gpu_launches']))"
