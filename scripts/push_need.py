"""CPU study: which share of today's NVLink pushes (every active contribution to every peer) is ever gathered by
the peer?  usage: python scripts/push_need.py <scale> <P>   (RMAT EF16, dealt round-robin ownership)"""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from memgraph_b200 import pagerank as pr
scale=int(sys.argv[1]); P=int(sys.argv[2])
n, m = 1<<scale, 16<<scale
s,t = pr.rmat_edges_host(scale, m)
indeg=np.bincount(t,minlength=n); outdeg=np.bincount(s,minlength=n)
order=np.lexsort((np.arange(n),-outdeg,-indeg))
pos=np.empty(n,dtype=np.int64); pos[order]=np.arange(n)
owner=pos%P            # dealt round-robin (owner of a vertex)
ou=owner[s]; ov=owner[t]
# distinct (u, owner(v)) pairs with owner(v) != owner(u)
key=s.astype(np.int64)*P+ov
key=key[ou!=ov]
pairs=np.unique(key).size
active=(outdeg>0).sum()
print(f"scale {scale} P={P}: active sources {active} ({active/n:.1%}); pushes today {active*(P-1)}; needed {pairs} = {pairs/(active*(P-1)):.1%}")
