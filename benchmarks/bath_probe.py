#!/usr/bin/env python
"""Ohmic-bath (DissipativeQuantumAnneal) pass probe on the 80x80 instance: ms per colour-pass launch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
R, P, S = int(os.environ.get("R", "4096")), int(os.environ.get("P", "64")), int(os.environ.get("S", "20"))
st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
st.init_random(1)
A, B = np.linspace(3.0, 1e-8, S), np.ones(S)
k = np.arange(1, P)
lut = 0.1 * (np.pi / (P * np.sin(np.pi * k / P))) ** 2  # qmc.pyx:162-163 kernel shape, alpha = 0.1
st.piqmc_sweeps_dissipative(A, B, 1, 1.0 / P, lut, seed=7)
inst.synchronize()
inst.timer_start()
st.piqmc_sweeps_dissipative(A, B, 1, 1.0 / P, lut, seed=7, sweep_offset=S)
ms = inst.timer_stop()
print("bath R=%d P=%d: %.1f us/launch, %.4g attempts/s" % (R, P, 1e3 * ms / (2 * S), R * S * P * inst.nspins / (ms * 1e-3)))
