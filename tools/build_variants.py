import sys, concurrent.futures as cf
sys.path.insert(0,'/root/repo')
from npbnn_b200 import build as B
V=dict(a.split('=',1) for a in sys.argv[1:])
V={k:[d for d in v.split(',') if d] for k,v in V.items()}
with cf.ThreadPoolExecutor(6) as ex:
    for n,o in zip(V, ex.map(lambda kv: B.build_variant(*kv), V.items())): print(n,o)
