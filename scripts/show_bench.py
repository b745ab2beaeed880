import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ('value', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'iters', d['config']['iterations'])
r = d['roofline']; print('roofline', round(r['achieved']), round(r['frac'], 3), 'share', round(r['share_of_matrix_kernel_time'], 3), {k: (v['launches'], round(v['ms'], 4), round(v['GBps'])) for k, v in r['per_mode'].items()})
tot = 0
for l in d['levels'][:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%9d %10d %3d %8.3f ms %7.0f GB/s  nnz/row %.1f" % (l['rows'], l['nnz'], l['launches'], l['ms'], l['GBps'], l['nnz'] / l['rows']))
print('sum matrix-kernel ms per solve', sum(l['ms'] for l in d['levels']))
for k, v in d.get('extra', {}).items():
    if 'unavailable' in v:
        print(k, v); continue
    print(k, {a: v[a] for a in ('value', 'iterations', 'true_relres', 'levels', 'host_setup_s', 'upload_s', 'gpu_launches_per_solve', 'operator_complexity') if a in v},
          'e2e', v['e2e']['value'], 'roofline', v['roofline'] and (round(v['roofline']['achieved']), round(v['roofline']['frac'], 3)), v.get('bsr_spmv'))
    for l in v.get('levels_table', [])[:8]:
        print('    ', l)
