"""Goldens for inputs beyond the bundled example, produced by THE REFERENCE (oracle/_ref/MIDASPOM.out) in the build
container: years with many missing cells and more than 32 distinct observation-compatible rows -- the cases the
reference's state tables (main_MIDASPOM.c:222-279) accept without limit.

    python tests/golden/make_golden_wide.py      ->  tests/golden/golden_wide.npz

  wide_a_obs / wide_a_post / wide_a_ltot   10 patches x 6 years, 6 missing cells in year 2 and 3 in year 4 (64 + 8 completions), -s 21
  wide_b_obs / wide_b_post / wide_b_ltot   8 patches x 45 years, complete surveys, 40 distinct rows, -s 11
Missing cells are kept out of year 0: with them the serial program reads an uninitialised Pold.
"""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402


def run_ref(obs, args):
    tmp = Path(tempfile.mkdtemp())
    (tmp / "in.txt").write_text("\n".join(" ".join(str(int(v)) for v in r) for r in obs) + "\n")
    out = subprocess.run([str(O.REF_BIN / "MIDASPOM.out"), "-i", str(tmp / "in.txt"), "-o", str(tmp / "p.txt")] + args,
                         check=True, capture_output=True, text=True).stdout
    post = np.array([[float(v) for v in line.split()] for line in (tmp / "p.txt").read_text().splitlines() if line.strip()])
    rows = re.findall(r"Year \d+: ([-\d ]+)\n", out.split("Input occupancy data:")[1].split("Number of possible")[0])
    echo = np.array([[int(v) for v in r.split()] for r in rows], dtype=np.int8)
    assert (echo == obs).all(), "the reference parsed a different table"
    return post, float(re.search(r"Total log-likelihood=([-\d.]+)", out).group(1))


def main():
    assert O.have_ref(), "oracle/_ref missing: run make -C oracle ref"
    rng = np.random.default_rng(20261019)
    g = {}
    # (a) a persistent metapopulation with two poorly surveyed years
    n, T = 10, 6
    z = np.zeros((T, n), dtype=np.int8)
    z[0] = rng.random(n) < 0.6
    for t in range(1, T):
        stay = z[t - 1] & (rng.random(n) > 0.3)
        z[t] = stay | (rng.random(n) < 0.25 * (stay.sum() > 0))
    obs = z.copy()
    obs[2, rng.choice(n, 6, replace=False)] = -1
    obs[4, rng.choice(n, 3, replace=False)] = -1
    g["wide_a_obs"] = obs
    g["wide_a_post"], g["wide_a_ltot"] = run_ref(obs, ["-m", "400", "-d", "100", "-s", "21"])
    # (b) a long complete record: more than 32 distinct rows
    n, T = 8, 45
    z = np.zeros((T, n), dtype=np.int8)
    z[0] = rng.random(n) < 0.6
    for t in range(1, T):
        stay = z[t - 1] & (rng.random(n) > 0.35)
        z[t] = stay | (rng.random(n) < 0.3)
    assert len({tuple(r) for r in z}) > 32
    g["wide_b_obs"] = z
    g["wide_b_post"], g["wide_b_ltot"] = run_ref(z, ["-m", "400", "-d", "100", "-s", "11"])
    np.savez_compressed(HERE / "golden_wide.npz", **g)
    print({k: np.shape(v) for k, v in g.items()}, "distinct rows (b):", len({tuple(r) for r in z}))


if __name__ == "__main__":
    main()
