import sys, json
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{"metric'):
        d = json.loads(line)
        r = d.get('roofline') or {}
        print(d['config']['workload'], '| value %.3e' % d['value'], '| ms/step', round(d['ms_per_step'], 3), '| e2e ms',
              round(d['e2e']['ms_per_step'], 3), '| roof', r.get('kernel'), round(r.get('frac', 0), 3), '|',
              {k: round(v, 3) for k, v in d['class_ms_per_step'].items()}, '| n_gpus', d['n_gpus'])
    elif line.startswith("{'mode'"):
        d = eval(line)
        print(d['mode'], 'filt', d.get('filt_rel_rms'), 'maxdiff>blk1', d['pcm_maxdiff_after_blk1'], 'exact', round(d['pcm_frac_exact'], 5))
    else:
        print(line[:220])
