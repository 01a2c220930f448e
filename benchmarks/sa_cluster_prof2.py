import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs
from bench import load_instance
nbs, _ = load_instance()
inst = mcs.Instance(nbs)
sched = np.linspace(3.0, 0.0, 1000)
os.environ["MCS_CLUSTER_VERBOSE"] = "1"
os.environ["MCS_CLUSTER"] = "1"
os.environ["MCS_CLUSTER_SIZE"] = "16"
for R, mcsteps, prof in ((1024, 1, None), (1024, 4, None), (1024, 4, "0"), (1024, 4, "50"), (1024, 4, "100"), (1024, 2, None), (2048, 1, None), (2048, 1, "0"), (2048, 1, "200")):
    os.environ.pop("MCS_CLUSTER_PROF", None)
    if prof is not None:
        os.environ["MCS_CLUSTER_PROF"] = prof
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    sc = sched[:1000 // mcsteps]
    st.sa_sweeps(sc, mcsteps, seed=3)
    inst.synchronize()
    for rep in range(2):
        inst.timer_start()
        st.sa_sweeps(sc, mcsteps, seed=3)
        ms = inst.timer_stop()
        print("R=%d mcsteps=%d prof=%s: %.3f ms, %.2f us per pass" % (R, mcsteps, prof, ms, ms / 2.0), flush=True)
    st.close()
