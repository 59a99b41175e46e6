import json, signal, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` must not produce BrokenPipe noise
j = json.load(open(sys.argv[1]))
for k in ('value', 'ms_per_step', 'dit_ms_per_step', 'frac_of_dense_gemm_roofline', 'gpu_launches', 'clocks', 'profiled_image_ms'):
    print(k, j.get(k))
r = j.get('roofline') or {}
print('roofline', {k: r.get(k) for k in ('achieved', 'frac', 'share_of_step', 'avg_launch_ms', 'traffic')})
print('e2e', j['e2e'])
print('cpu', (j.get('cpu_baseline') or {}).get('value'), (j.get('cpu_baseline') or {}).get('cores'))
for k, v in j['kernels'].items():
    if v['launches']:
        extra = f"  {v['flops']/v['ms']/1e9:8.1f} TFLOP/s" if v['flops'] else f"  {v['bytes']/v['ms']/1e6:8.1f} GB/s"
        print(f"{k:20s} ms={v['ms']:8.1f} n={v['launches']:6d} avg_us={1e3*v['ms']/v['launches']:8.1f}{extra}")
