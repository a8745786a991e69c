import numpy as np, sys
sys.path.insert(0,'.')
from simc_gfortran_b200 import Simc, load_optics_fixture
from tests.oracle_lib import Oracle, transport_inputs
orc=Oracle(); sim=Simc(mode="strict")
for arm in (1,5):
    t=load_optics_fixture(arm); orc.set_optics(t); sim.set_optics(t)
    inp=transport_inputs(arm,20000,seed=99+arm)
    for ms,wcs in ((True,True),(True,False),(False,False)):
        ro,rf=orc.transport_batch(arm,inp,seed=4242,ms=ms,wcs=wcs)
        o,f=sim.transport_batch(arm,inp,4242,ms_flag=ms,wcs_flag=wcs)
        bad=np.nonzero(o[11]!=ro[11])[0]
        print(arm,ms,wcs,"flags equal",np.array_equal(f,rf),"draw mismatches",len(bad))
        for i in bad[:10]: print("  row",i,"flag",f[i],"draws",o[11][i],ro[11][i])
        import collections
        print("  by flag:",collections.Counter(f[bad].tolist()))
