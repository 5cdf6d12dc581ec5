import json, os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
for r in ('1', '2', '1', '2'):
    os.environ['B200OCR_CONFIG3_REPLICAS'] = r
    out = bench.bench_config3(torch.device('cuda', 0), 1414.1, False)
    print(r, round(out['value']), out['ms_per_256_lines'], out['forward_only'])
