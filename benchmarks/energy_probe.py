import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import montecarlosolvers_b200 as mcs
from bench import load_instance
nbs,_=load_instance()
inst=mcs.Instance(nbs)
for R in (4096, 512):
    st=mcs.State(inst, mcs._lib.KIND_PIQMC, R, 64)
    st.init_random(1)
    st.energies()
    for rep in range(2):
        inst.timer_start(); e=st.energies(); ms=inst.timer_stop()
        print("R=%d energies(): %.3f ms"%(R,ms))
    st.close()
