"""CPU checks of two pieces of host/device-shared spatial logic (no GPU needed):

* `hilbert3` (skid_b200/csrc/common.cuh), the sort key of every tree and of the mover tiles: compiled for the host
  with g++ and checked to be a bijection whose consecutive indices are face neighbours (what makes a run of 32
  consecutive sorted points one connected blob);
* the bounds by which `k_link_cells` (skid_b200/csrc/fof.cu) decides a pair of FoF cells from the bounding boxes of
  their movers: restated in numpy and compared with the brute-force min-image distances of random point sets, with
  and without periodic wrap."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

SRC = r"""
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
int main(int argc, char **argv)
{
	if (argc > 1) { // keys of the coordinates given on stdin (x y z per line), bits = argv[1]
		const int bits = atoi(argv[1]);
		unsigned x, y, z;
		while (scanf("%u %u %u", &x, &y, &z) == 3) printf("%llu\n", (unsigned long long)hilbert3(x, y, z, bits));
		return 0;
	}
	for (int bits = 1; bits <= 6; ++bits) {
		const int n = 1 << bits;
		const size_t tot = (size_t)n * n * n;
		std::vector<int> cx(tot), cy(tot), cz(tot);
		std::vector<char> seen(tot, 0);
		for (int x = 0; x < n; ++x)
			for (int y = 0; y < n; ++y)
				for (int z = 0; z < n; ++z) {
					const uint64_t k = hilbert3(x, y, z, bits);
					if (k >= tot || seen[k]) { printf("bits %d: not a bijection\n", bits); return 1; }
					seen[k] = 1; cx[k] = x; cy[k] = y; cz[k] = z;
				}
		for (size_t k = 1; k < tot; ++k)
			if (abs(cx[k] - cx[k - 1]) + abs(cy[k] - cy[k - 1]) + abs(cz[k] - cz[k - 1]) != 1) {
				printf("bits %d: indices %zu and %zu are not face neighbours\n", bits, k - 1, k); return 1;
			}
	}
	printf("OK\n");
	return 0;
}
"""


def _build(tmp_path):
    src = tmp_path / "hilbert_check.cpp"
    exe = tmp_path / "hilbert_check"
    src.write_text(SRC)
    inc = [os.path.join(ROOT, "skid_b200", "csrc"), "/usr/local/cuda/include"]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w"] + [f"-I{i}" for i in inc] + ["-o", str(exe), str(src)], check=True)
    return str(exe)


def test_hilbert_key_is_a_face_connected_bijection(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr


def test_hilbert_runs_are_compact(tmp_path):
    """Runs of 32 consecutive keys on a 16^3 lattice: the bounding box of a Hilbert run holds at most 4x its 32 cells
    (a run along the Z curve can span the whole lattice)."""
    exe = _build(tmp_path)
    n = 16
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    r = subprocess.run([exe, "4"], input="\n".join(f"{a} {b} {c}" for a, b, c in g), capture_output=True, text=True)
    keys = np.array(r.stdout.split(), np.int64)
    order = np.argsort(keys)
    runs = g[order].reshape(-1, 32, 3)
    vol = np.prod(runs.max(1) - runs.min(1) + 1, axis=1)
    assert vol.max() <= 128, vol.max()


def _cell_bounds(alo, ahi, blo, bhi, L):
    """numpy mirror of the box test in k_link_cells: (lower bound, upper bound) of the squared min-image distance
    between any point of box A and any point of box B."""
    direct = np.maximum(np.maximum(blo - ahi, alo - bhi), 0.0)
    span = np.maximum(ahi, bhi) - np.minimum(alo, blo)
    wrapped = np.maximum(L - span, 0.0)
    gmin = np.minimum(direct, wrapped)
    sep = np.maximum(ahi - blo, bhi - alo)
    return float(np.sum(gmin * gmin)), float(np.sum(sep * sep))


def test_fof_cell_box_bounds_against_brute_force():
    rng = np.random.default_rng(5)
    for trial in range(4000):
        periodic = trial % 2 == 0
        L = np.full(3, 1.0) if periodic else np.full(3, 3.0e38)
        ca = rng.random(3) - 0.5
        # B near A, or near A's periodic image across a face
        cb = ca + (rng.random(3) - 0.5) * 0.02
        if periodic and trial % 4 == 0:
            cb[rng.integers(3)] += rng.choice([-1.0, 1.0]) * (1.0 - rng.random() * 0.01)
        a = ca + (rng.random((rng.integers(1, 12), 3)) - 0.5) * 0.006
        b = cb + (rng.random((rng.integers(1, 12), 3)) - 0.5) * 0.006
        d = a[:, None, :] - b[None, :, :]
        if periodic:
            d = d - np.round(d)          # min image (points may lie outside the box: |d| <= 1.5 L)
        d2 = np.sum(d * d, axis=2)
        lo2, hi2 = _cell_bounds(a.min(0), a.max(0), b.min(0), b.max(0), L)
        assert lo2 <= d2.min() * (1 + 1e-9) + 1e-18, (trial, lo2, d2.min())
        assert hi2 >= d2.max() * (1 - 1e-9), (trial, hi2, d2.max())
