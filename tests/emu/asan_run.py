import sys, ctypes
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests','emu')); sys.path.insert(0,os.path.join(R,'tests'))
import pyemu
pyemu._LIB='/tmp/libemu_asan.so'
pyemu.build=lambda force=False: pyemu._LIB
from oracle import pyoracle as po
from common import *
import numpy as np
st=po.default_settings()
g=load_golden()
idx=[i for i,n in enumerate(g['names']) if n.startswith(('uf20-01','uf50-01','uuf50-01','2bitcomp'))]
ioff,coff,lits=batch_of([golden_instance(g,i) for i in idx])
for kind in (0,1):
    for tier in (-1,0):
        res,_=pyemu.solve_batch(ioff,coff,lits,st,kind=kind,tier_nv=tier,model_stride=int(max(g['n_vars'][i] for i in idx)))
        print('kind',kind,'tier',tier,[ (r.status,r.verdict) for r in res], flush=True)
# tiny arena: retry ladder paths
ioff,coff,lits=synth_batch(po,SEED2+7000,2,50,213)
res,_=pyemu.solve_batch(ioff,coff,lits,st,kind=0,arena_words=1200,wpool_entries=700)
print('retry',[ (r.status,r.attempts) for r in res])
off,l2=golden_instance(g,list(g['names']).index('uuf50-01.cnf.gz'))
st2=po.default_settings(); st2.restart_first=6.0
res,_=pyemu.portfolio(off,l2,st2,4,share=3|4,model_stride=50)
print('portfolio',[ (r.status,r.verdict,r.imported) for r in res])
