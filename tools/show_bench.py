#!/usr/bin/env python
"""Pretty-prints a bench.py JSON line and (optionally) an ncu launch-list CSV."""
import collections
import csv
import json
import sys


def show_bench(path):
    d = json.load(open(path))
    print('value %.3e %s  ms_step %.4f  launches %d  n_gpus %d' % (d['value'], d['unit'], d['ms_per_step'], d['gpu_launches'], d['n_gpus']))
    e = d['e2e']; print('e2e %.3e ms %.4f' % (e['value'], e['ms_per_step']))
    r = d['roofline']; print('K2 ms %.4f achieved %.1f TF frac %.3f share %.2f exact=%s fb=%s' % (r['kernel_ms'], r['achieved'], r['frac'], r['kernel_share_of_step'], r['exact_integer_mode'], r['exact_fallback_rows']))
    if r.get('in_chain'): print('   in chain: K2 %.2f us, %.1f TF, frac %.3f' % (r['in_chain']['kernel_ms'] * 1e3, r['in_chain']['achieved'], r['in_chain']['frac']))
    if d['e2e'].get('ms_per_step_blocks'): print('   e2e blocks (ms):', ' '.join('%.3f' % x for x in d['e2e']['ms_per_step_blocks']))
    print('clocks', d['clocks'])
    if d.get('cpu_baseline'):
        print('cpu %.3e cores %d' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
    s = d.get('secondary')
    if s:
        print('ransac %.3e hyp/s ms %.2f e2e %.3e k7_ms %.2f frac %.3f inl %s' % (s['value'], s['ms_per_step'], s['e2e']['value'], s['roofline']['kernel_ms'], s['roofline']['frac'], s['config']['winner_inliers']), s.get('cpu_baseline', {}).get('value'))
    x = (d.get('extra') or {}).get('hamming')
    if x:
        for k in ('popc', 'tensor'):
            if k in x:
                y = x[k]
                print('hamming[%s] step ms %.3f kernel ms %.3f pairs/s(kernel) %.3e frac %.3f mutual %d' % (k, y['ms_per_step'], y['kernel_ms'], y['pairs_per_s_knn_kernel'], y['roofline']['frac'], y['mutual_matches']))
        if 'popc' not in x:
            print('hamming step ms %.3f kernel ms %.3f pairs/s(kernel) %.3e frac %.3f mutual %d' % (x['ms_per_step'], x['kernel_ms'], x['pairs_per_s_knn_kernel'], x['roofline']['frac'], x['mutual_matches']))
    sf = (d.get('extra') or {}).get('l2_general_floats')
    if sf:
        print('surf-like floats: step %.1f us %.3e pairs/s fallback rows %d' % (sf['ms_per_step'] * 1e3, sf['pairs_per_s'], sf['exact_fallback_rows']))
    c5 = (d.get('extra') or {}).get('cfg5')
    if c5:
        print('cfg5 %.1f image pairs/s (%.3f ms/pair) last %s' % (c5['image_pairs_per_s'], c5['ms_per_pair'], c5['last_pair']))
        if c5.get('e2e'):
            print('   from host buffers: %.1f image pairs/s (%.3f ms/pair)' % (c5['e2e']['value'], c5['e2e']['ms_per_pair']))
    if d.get('summary'):
        print('parity', d['summary'].get('parity'))


def show_launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
            continue
        agg.setdefault(r[kn][:70], []).append(float(r[mv].replace(',', '')))
    tot = sum(sum(v) / len(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k:70s} n={len(v):4d} avg_us={sum(v)/len(v)/1000:8.2f} share={sum(v)/len(v)/tot:5.2f}")
    print('sum of avgs us %.2f' % (tot / 1000))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        (show_launches if p.endswith('.csv') else show_bench)(p)
