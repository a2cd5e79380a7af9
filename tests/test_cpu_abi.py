"""CPU-only checks of the boundary and host logic: the C-ABI library loads and exports every symbol
include/skidgpu.h declares, it fails loudly without a GPU (no CPU fallback), struct layouts match the
reference's PINIT/PGROUP, TIPSY I/O round-trips, the host driver keeps the reference's usage() contract,
and the golden fixtures are self-consistent."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import refdump
from skid_b200 import api, synth, tipsy


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-j8", "-C", ROOT, "lib"], check=True, capture_output=True)
    subprocess.run(["make", "-C", ROOT, "host"], check=True, capture_output=True)
    return True


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "skidgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(skidgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(api.EXPORTS), set(declared) ^ set(api.EXPORTS)
    lib = api.load_library()
    for s in declared:
        assert hasattr(lib, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    for s in declared:
        assert re.search(rf"\bT {s}\b", nm), s


def test_struct_layouts():
    # PINIT 48 B, PGROUP 68 B (measured sizeof of kd.h:27-55)
    assert tipsy.PINIT_DTYPE.itemsize == 48 and tipsy.PINIT_DTYPE.fields["iOrder"][1] == 44
    assert tipsy.PGROUP_DTYPE.itemsize == 68 and tipsy.PGROUP_DTYPE.fields["nMembers"][1] == 56
    # skidgpu_stat_row (include/skidgpu.h) == the ctypes mirror == the oracle's row: int + 10 floats
    from oracle import orc
    assert api.STAT_ROW_DTYPE.itemsize == 44 and api.STAT_ROW_DTYPE == orc.STAT_ROW_DTYPE
    hdr = open(os.path.join(ROOT, "include", "skidgpu.h")).read()
    body = re.search(r"typedef struct \{([^}]*)\} skidgpu_stat_row;", hdr).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [t for t in re.findall(r"\b(nMembers|f[A-Za-z0-9]+)\b", body) if t != "float"]
    assert names == list(api.STAT_ROW_DTYPE.names), names


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.SkidError, match="no CPU fallback"):
        api.SkidGPU()


def test_product_never_imports_oracle():
    bad = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|skid_oracle|orc_[a-z_]+\s*\()")
    for top in ("skid_b200", "host", "include"):
        for root, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                    text = open(os.path.join(root, f)).read()
                    assert not bad.search(text), (f, "product code must not use the oracle")


def test_tipsy_roundtrip(tmp_path):
    snap = synth.make_box(2048, seed=5, kind="gasdark")
    for std in (True, False):
        f = tmp_path / ("a.std" if std else "a.bin")
        gas, dark, star = tipsy.pinit_to_records(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"])
        tipsy.write_tipsy(f, 1.0, gas, dark, star, standard=std)
        back = tipsy.read_tipsy(f, standard=std)
        assert (back["nGas"], back["nDark"], back["nStar"]) == (snap["nGas"], snap["nDark"], 0)
        for k in ("r", "v", "fMass", "fSoft", "fTemp"):
            assert np.array_equal(back["pinit"][k], snap["pinit"][k]), k


def test_synth_box_properties():
    s = synth.make_box(1 << 14, seed=1)
    p = s["pinit"]
    assert p["r"].min() > -0.5 and p["r"].max() <= 0.5
    assert len(np.unique(p["r"], axis=0)) == len(p)          # no duplicate positions
    assert abs(p["fMass"].sum() - 1.0) < 1e-4
    assert s["flags"]["tau"] == pytest.approx(0.0288 * (1 << 14) ** (-1 / 3), rel=1e-6)


def test_host_driver_usage_contract(built):
    exe = os.path.join(ROOT, "host", "skid")
    # no -tau -> usage, exit 1 (main.c:339); unknown flag -> usage, exit 1 (main.c:334)
    for args in ([], ["-tau", "1", "-bogus"], ["-tau"]):
        r = subprocess.run([exe] + args, stdin=subprocess.DEVNULL, capture_output=True, text=True)
        assert r.returncode == 1 and r.stderr.startswith("USAGE:"), args


def test_host_driver_accepts_reference_flags(built, tmp_path):
    """Every flag of the reference's main.c:140-335 parses; without a GPU the run then stops at
    skidgpu_create with the 'no CPU fallback' message (never at usage())."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    snap = synth.make_box(512, seed=2)
    f = tmp_path / "in.std"
    synth.write_std(snap, str(f))
    flags = ["-std", "-tau", "9e-4", "-z", "0.5", "-O", "0.3", "-Lambda", "0.7", "-Q", "0", "-G", "1", "-H", "2.8944",
             "-s", "32", "-d", "170", "-t", "3e4", "-M", "1", "-fic", "-cvg", "4e-4", "-scoop", "2e-3", "-m", "8",
             "-maxgroup", "100000", "-nu", "-gd", "-go", "-spline", "-plummer", "-e", "1e-3", "-p", "1", "-c", "0",
             "-cx", "0", "-cy", "0", "-cz", "0", "-o", str(tmp_path / "out"), "-ray", "-den", "-stats", "-diag", "-nsp"]
    with open(f, "rb") as fin:
        r = subprocess.run([os.path.join(ROOT, "host", "skid")] + flags, stdin=fin, capture_output=True, text=True)
    assert r.returncode == 1
    assert "USAGE" not in r.stderr and "no CPU fallback" in r.stderr
    assert "nDark:512 nGas:0 nStar:0" in r.stdout


def test_golden_fixture_consistency(demo_golden, demo_input):
    p = demo_input[0]
    assert len(p) == 32768 and demo_input[2] == 32768
    assert str(demo_golden["md5_grp"]) == "bbef8709daa0e9d5525981137cf1b02b"
    assert str(demo_golden["md5_den"]) == "fc56c1b53a179c87654888d196cf5343"
    assert int(demo_golden["nGroup"]) == 68 and int(demo_golden["nUnbound"]) == 4134
    assert int(demo_golden["grp"].max()) == 68
    assert tuple(demo_golden["ittr"][0]) == (0, 12190, 20891) and tuple(demo_golden["ittr"][-1]) == (37, 0, 8994)
    lab = refdump.canonical_labels(demo_golden["grp"])
    assert lab.max() == 68 and np.array_equal(lab == 0, demo_golden["grp"] == 0)


def test_cosmology_scalar():
    # Einstein-de Sitter: H(a) = H0 a^-3/2 ; flat Lambda: H(1) = H0
    assert api.csmExp2Hub(0.5, 2.0, 1.0, 0.0) == pytest.approx(2.0 * 0.5 ** -1.5)
    assert api.csmExp2Hub(1.0, 2.8944, 0.3, 0.7) == pytest.approx(2.8944)
