import sys, json
for l in open(sys.argv[1]):
    if not l.startswith('{'):
        print(l, end=''); continue
    d = json.loads(l)
    print(d['shape'], d['grid'], 'st', d['stages'], 'avg', d['avg_us_per_launch'], 'spread', d['cta_start_spread_us'], 'span', d['kernel_span_us'],
          ' '.join(f"{k.split('_cyc')[0]}={d[k][0]}" for k in d if 'cyc' in k))
