"""CPU-side checks: the C-ABI library loads and exports every symbol include/mcs_b200.h declares (no compute
calls without a GPU), fails loudly without a device, and the host-side helpers (tools, sharding) are right."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import oracle as orc
from tests import instances as inst

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mcs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import montecarlosolvers_b200 as m
    L = m._lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libmcs_b200.so does not export %s" % n
    assert sorted(m._lib.SIGNATURES) == names  # the ctypes table mirrors the header one to one
    assert L.mcs_abi_version() == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "montecarlosolvers_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                # comments may NAME the oracle (it defines the energy convention); code may not reach it
                for pat in (r"^\s*(from|import)\s+oracle", r"import_module\([\"']oracle", r"#include\s+[\"<][^\n]*oracle",
                            r"libmcs_oracle", r"oracle/_ref", r"oracle\.(py|c)\b"):
                    assert not re.search(pat, txt, flags=re.M), (f, pat)


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_fails_loudly_without_a_device():
    import montecarlosolvers_b200 as m
    _, nbs = inst.torus(4, seed=1)
    assert m.device_count() == 0
    with pytest.raises(m.McsError, match="no CPU fallback"):
        m.Instance(nbs)
    c = np.ones((16, 4), dtype=np.int64)
    with pytest.raises(m.McsError):
        m.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.1, c, nbs, 1)
    assert np.all(c == 1)


def test_argument_validation_before_any_device_work():
    import montecarlosolvers_b200 as m
    _, nbs = inst.torus(4, seed=1)
    c = np.ones((16, 4), dtype=np.int64)
    with pytest.raises(ValueError):
        m.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.1, c.astype(float), nbs, 1)  # dtype mismatch
    with pytest.raises(ValueError):
        m.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.1, c, nbs.astype(np.float32), 1)
    with pytest.raises(ValueError):
        m.sa.Anneal(np.ones((2, 2)), 1, c[:, 0].copy(), nbs)  # wrong ndim
    with pytest.raises(TypeError):
        m.sa.Anneal(np.ones(2), 1, [1, -1], nbs)
    with pytest.raises(ValueError):
        m.svmc.SpinVectorMonteCarlo(np.ones(2), np.ones(2), 1, 0.1, np.ones(16, dtype=np.float32), nbs)


def test_tools_match_the_reference_format():
    import montecarlosolvers_b200 as m
    J, nbs = inst.random_graph(30, 70, seed=9)
    assert np.array_equal(m.tools.GenerateNeighbors(30, J, nbs.shape[1]), nbs)
    for seed in range(3):
        s = inst.random_spins(30, seed)
        assert abs(m.tools.ClassicalIsingEnergy(s, J) - orc.ClassicalIsingEnergy(s, J)) < 1e-10
        assert abs(m.tools.ClassicalIsingEnergy(s, J) - orc.ising_energy(s, nbs)) < 1e-10
    with pytest.raises(ValueError):
        m.tools.GenerateNeighbors(30, J, 1)  # the reference overflows silently (no bounds check)
    assert m.tools.bits2spins([0, 1]) == [1, -1] and m.tools.spins2bits([1, -1]) == [0, 1]
    Jd = sps.dok_matrix((3, 3))
    Jd[0, 1] = 2.0
    Jd[2, 2] = -0.5
    t = m.tools.GenerateNeighbors(3, Jd, 2)
    assert t[0, 0].tolist() == [1.0, 2.0] and t[1, 0].tolist() == [0.0, 2.0] and t[2, 0].tolist() == [2.0, -0.5]


def test_shard_covers_replicas_exactly_once():
    from montecarlosolvers_b200 import parallel
    for R in (1, 7, 4096, 1000):
        for world in (1, 2, 3, 8):
            b = [parallel.shard(R, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == R
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
            a = [parallel.shard_aligned(R, r, world) for r in range(world)]
            assert a[0][0] == 0 and a[-1][1] == R and all(l % 32 == 0 for l, h in a if h > l)
            assert sum(h - l for l, h in a) == R


def test_build_stamp_is_a_content_hash_that_survives_copying_the_tree(tmp_path):
    """The GPU box runs a COPY of the tree at another path with fresh mtimes: the library must count as up to date
    there (every rank of a torchrun launch would otherwise rebuild it at once), and must stop counting as up to date
    as soon as a source changes."""
    import shutil
    import subprocess
    import sys
    from montecarlosolvers_b200 import build
    if not os.path.isfile(build.OUT):
        pytest.skip("library not built")
    assert build.up_to_date()
    root = os.path.dirname(os.path.dirname(os.path.abspath(build.__file__)))
    dst = tmp_path / "copy"
    (dst / "montecarlosolvers_b200").mkdir(parents=True)
    shutil.copytree(os.path.join(root, "include"), dst / "include")
    for name in os.listdir(os.path.join(root, "montecarlosolvers_b200")):
        src = os.path.join(root, "montecarlosolvers_b200", name)
        if name == "csrc":
            shutil.copytree(src, dst / "montecarlosolvers_b200" / "csrc")
        elif os.path.isfile(src) and (name.endswith(".py") or name.endswith(".srchash")):
            shutil.copy(src, dst / "montecarlosolvers_b200" / name)
    (dst / "montecarlosolvers_b200" / "libmcs_b200.so").write_bytes(b"")  # presence is enough for the check
    code = "from montecarlosolvers_b200 import build; print(build.up_to_date())"
    run = lambda: subprocess.run([sys.executable, "-c", code], cwd=str(dst), stdout=subprocess.PIPE, text=True,
                                 env=dict(os.environ, PYTHONPATH=str(dst))).stdout.strip()
    assert run() == "True"
    with open(dst / "montecarlosolvers_b200" / "csrc" / "mcs_sa.cu", "a") as f:
        f.write("// touched\n")
    assert run() == "False"
