import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        if line: print('  [log]', line[:200])
        continue
    d=json.loads(line)
    k=d['roofline']['kernels']
    print('%s: %.1f it/s  %.3f ms | fwd %.3f (%.2f) col %.3f (%.2f) prox %.3f (%.2f) sc %.4f | e2e %.1f' % (sys.argv[1] if len(sys.argv)>1 else '', d['value'], d['ms_per_step'], k['k_row_fwd']['ms'], k['k_row_fwd']['frac'], k['k_col']['ms'], k['k_col']['frac'], k['k_row_inv_prox']['ms'], k['k_row_inv_prox']['frac'], k['k_admm_scalars']['ms'], d['e2e']['value']), ' iter_frac %.3f' % d['roofline']['iteration_frac'])
