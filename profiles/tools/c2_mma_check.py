import sys, json, numpy as np, torch
sys.path.insert(0, '.')
from genes2genes_b200 import synthetic
sys.path.insert(0, 'profiles/tools')
import k1m_check as kc
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
kc.case(1371, 20327, 17176, 14, 0.9, 1385, dev, flush, n_oracle=1371)
kc.case(994, 3157, 890, 13, 0.9, 1007, dev, flush, n_oracle=994)
