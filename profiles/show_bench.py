import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        l = l.strip()
        if l.startswith('{'):
            d = json.loads(l)
            ks = d.pop('kernels', None)
            print(f, 'value %.3fM frames/s  ms/step %.3f  e2e %.3fM  launches %s' % (d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6, d.get('gpu_launches')))
            print('   roofline', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'ms_per_launch')}, 'halves', {k: round(v, 3) for k, v in d['halves'].items() if k.endswith('ms')})
            if 'cpu_baseline' in d: print('   cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
            for k in ks or []:
                print('   %-28s %7.3f ms  %7.1f GB/s frac %.3f launches %d' % (k['name'], k['ms_per_step'], k['gbs'], k['frac'], k['launches_per_step']))
