"""Prints the fields of a bench.py JSON line that matter when comparing runs."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith('{')][-1])
    except Exception as exc:
        print(path, 'unreadable:', exc)
        continue
    print('%s: n_gpus %s  ms/step %.2f  value %.3e  e2e %.2f ms (%.3e)  filter frac %.3f  fallback %s' % (
        path, d.get('n_gpus'), d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'],
        d['roofline']['frac'], d.get('details', d['config']).get('fallback_rows_last_step')))
    ph = d.get('phases_ms', {})
    for key in ('rank0', 'max_over_ranks', 'mean_over_ranks'):
        if key in ph and isinstance(ph[key], dict):
            print('   %-16s %s' % (key, '  '.join('%s %.2f' % kv for kv in ph[key].items())))
    if 'unsharded_share_of_step' in ph:
        print('   unsharded share of step:', ph['unsharded_share_of_step'])
    print('   parity', {k: v for k, v in d.get('parity', {}).items() if k != 'against'})
    print('   clocks', d.get('clocks'))
