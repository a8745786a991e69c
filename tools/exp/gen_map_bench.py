#!/usr/bin/env python
"""Experiment: straight-line code for one COSY class vs the record interpreter's cost model.
   python tools/exp/gen_map_bench.py hms 5 > /tmp/map_bench.cu"""
import sys, numpy as np
name, klasses = sys.argv[1], [int(x) for x in sys.argv[2].split(',')]
z = np.load(f'tests/golden/optics_{name}.npz')
cs, E, C = z['class_start'], z['fwd_expon'], z['fwd_coeff']
V = ['x', 't', 'y', 'p', 'd']
def pw(v, e):
    return v if e == 1 else f'{v}{e}'
import os
CONSTTAB = os.environ.get('CONSTTAB') == '1'
ctab = []
out = ['#include <cstdio>', '#include <cuda_runtime.h>', '@@CT@@']
for k in klasses:
    e = E[cs[k-1]:cs[k]]; c = C[cs[k-1]:cs[k]]
    out.append(f'__global__ void __launch_bounds__(128) map_{k}(const double* __restrict__ in, double* __restrict__ o, long long n) {{')
    out.append('  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {')
    for j, v in enumerate(V):
        out.append(f'  const double {v} = in[{j} * n + i];')
        mx = int(e[:, j].max())
        # libgcc __powidf2 association
        if mx >= 2: out.append(f'  const double {v}2 = {v} * {v};')
        if mx >= 3: out.append(f'  const double {v}3 = {v} * {v}2;')
        if mx >= 4: out.append(f'  const double {v}4 = {v}2 * {v}2;')
        if mx >= 5: out.append(f'  const double {v}5 = {v} * {v}4;')
        if mx >= 6: out.append(f'  const double {v}6 = {v}2 * {v}4;')
    out.append('  double s0 = 0., s1 = 0., s2 = 0., s3 = 0., s4 = 0.;')
    for r in range(len(e)):
        if not (c[r] != 0).any(): continue
        fac = [pw(V[j], int(e[r, j])) for j in range(5) if e[r, j] > 0]
        if not fac: m = '1.0'
        else:
            m = fac[0]
            for f in fac[1:]: m = f'({m} * {f})'
        out.append(f'  {{ const double m = {m};')
        for q in range(5):
            if c[r, q] != 0:
                if CONSTTAB:
                    out.append(f'    s{q} = s{q} + m * CT[{len(ctab)}];'); ctab.append(float(c[r, q]).hex())
                else:
                    out.append(f'    s{q} = s{q} + m * {float(c[r, q]).hex()};')
        out.append('  }')
    out.append('  o[0 * n + i] = s0; o[1 * n + i] = s1; o[2 * n + i] = s2; o[3 * n + i] = s3; o[4 * n + i] = s4;')
    out.append('  }\n}')
out.append('''
int main() {
  const long long n = 1 << 22;
  double *in, *o; cudaMalloc(&in, 5 * n * 8); cudaMalloc(&o, 5 * n * 8);
  double* h = (double*)malloc(5 * n * 8);
  for (long long i = 0; i < 5 * n; ++i) h[i] = ((i * 2654435761u) % 1000) * 1e-3 - 0.5;
  cudaMemcpy(in, h, 5 * n * 8, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;''')
for k in klasses:
    nt = int(((C[cs[k-1]:cs[k]] != 0).any(axis=1)).sum())
    out.append(f'''  for (int rep = 0; rep < 3; ++rep) {{ cudaEventRecord(a); map_{k}<<<148 * 8, 128>>>(in, o, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); }}
  printf("class {k}: {nt} terms, %.3f ms per 4M events, %.2f cycles per term per warp per SM\\n", ms, ms * 1e-3 * 1.965e9 * 148 / ((double)n / 32 * {nt}));''')
out.append('  return 0;\n}')
txt='\n'.join(out)
print(txt.replace('@@CT@@', ('__constant__ double CT[] = {' + ', '.join(ctab) + '};') if CONSTTAB else ''))
