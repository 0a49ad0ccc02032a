"""The C drivers (driver/midaspom, driver/midaspom_future) keep the reference's flags and file
formats and call the engine through the C ABI only.  Compared with MIDASPOM.out /
MIDASPOM_future.out outputs recorded in tests/golden/golden.npz (make_golden.py)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
EXAMPLE = ROOT / "tests" / "golden" / "occupancies_example.txt"


def build_drivers():
    subprocess.run(["make", "-C", str(ROOT / "driver")], check=True, capture_output=True)


def read_table(path):
    return np.array([[float(v) for v in line.split()] for line in Path(path).read_text().splitlines() if line.strip()])


def run(exe, args):
    res = subprocess.run([str(ROOT / "driver" / exe)] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def test_driver_reproduces_posterior_file_of_run_examples(golden, tmp_path):
    """run_examples.sh:8 -- same flags, same ragged input file, same output format."""
    build_drivers()
    out = run("midaspom", ["-m", 400, "-d", 100, "-i", EXAMPLE, "-o", tmp_path / "posterior.txt"])
    assert "Number of habitat patches: 8\nNumber of sampled years: 7\n" in out         # stream-order parse of the ragged file
    assert "Year 1: 0 -1 1 0 1 0 1 0 \n" in out and "Year 2: 1 0 1 1 0 1 0 1 \n" in out
    assert "Year 0: 1\nYear 1: 2\nYear 2: 1\nYear 3: 2\nYear 4: 1\nYear 5: 2\nYear 6: 1\n" in out
    assert "Number of states to compute: 256\n" in out
    assert "Total log-likelihood=-39.34251\n" in out
    raw = (tmp_path / "posterior.txt").read_text()
    assert raw.count("\n") == 101 and all(len(l.split("\t")) == 102 for l in raw.splitlines())   # "%.20lf\t" x 101 per row
    np.testing.assert_allclose(read_table(tmp_path / "posterior.txt"), golden["post_default"], rtol=1e-9, atol=1.1e-20)
    out = run("midaspom", ["-m", 400, "-d", 100, "-s", 3, "-l", 0.3, "-u", 0.7, "-i", EXAMPLE, "-o", tmp_path / "p3.txt"])
    assert "Total log-likelihood=-40.29644\n" in out
    np.testing.assert_allclose(read_table(tmp_path / "p3.txt"), golden["post_s3"], rtol=1e-9)


def test_driver_mcmc_mode_writes_the_same_format(golden, tmp_path):
    """-n switches the grid loops for chains; the (e, c) histogram goes onto the same s x s grid."""
    build_drivers()
    out = run("midaspom", ["-m", 400, "-d", 100, "-s", 21, "-n", 9000, "-c", 8, "-i", EXAMPLE, "-o", tmp_path / "pm.txt",
                           "-t", tmp_path / "draws.txt"])
    tab = read_table(tmp_path / "pm.txt")
    assert tab.shape == (21, 21)
    w1 = np.ones(21); w1[0] = w1[-1] = 0.5
    assert abs((np.outer(w1, w1) * tab).sum() * 0.05 ** 2 - 1.0) < 1e-9                # trapezoid mass 1, like the reference's table
    grid = np.linspace(0, 1, 21)
    pm = (np.outer(w1, w1) * tab); pm /= pm.sum()
    exact = np.outer(np.r_[0.5, np.ones(99), 0.5], np.r_[0.5, np.ones(99), 0.5]) * golden["post_default"]; exact /= exact.sum()
    g101 = np.linspace(0, 1, 101)
    assert abs((pm.sum(1) * grid).sum() - (exact.sum(1) * g101).sum()) < 0.01         # E[e]
    assert abs((pm.sum(0) * grid).sum() - (exact.sum(0) * g101).sum()) < 0.015        # E[c]
    assert len((tmp_path / "draws.txt").read_text().splitlines()) == 9000 * 8 + 1


@pytest.mark.parametrize("tag,extra", [("nomgmt", []), ("source", ["-S", 1, "-s", 500])])
def test_future_driver_matches_reference_runs(golden, tmp_path, tag, extra):
    """run_examples.sh:17,20 -- extinct counts per future year vs 6 runs of MIDASPOM_future.out."""
    build_drivers()
    run("midaspom", ["-m", 400, "-d", 100, "-i", EXAMPLE, "-o", tmp_path / "posterior.txt"])
    run("midaspom_future", ["-a", 50, "-m", 400, "-d", 100, "-i", EXAMPLE, "-q", tmp_path / "posterior.txt",
                            "-o", tmp_path / "f.txt", "-r", 5] + extra)
    got = np.array([int(v) for v in (tmp_path / "f.txt").read_text().split()])
    runs = golden[f"future_{tag}_runs"]
    assert got.shape == (50,)
    mu = runs.mean(axis=0)
    p = np.clip(mu / 10000.0, 1e-4, 1 - 1e-4)
    sd = np.sqrt(10000 * p * (1 - p) * (1 + 1 / len(runs)))
    assert (np.abs(got - mu) < 5 * sd + 3).all()
    if not extra:
        assert (np.diff(got) >= 0).all()                                               # extinction is absorbing without a source


def test_variant_drivers_and_hypothesis_numbers(golden, tmp_path):
    """run_examples.sh:11,14 with the reference's flags; then the AIC / Bayes-factor numbers of
    Rscript/hypothesis_test.R:29-46 computed from the two files."""
    build_drivers()
    run("midaspom_dieoff", ["-a", 10, "-e", 0.71, "-c", 0.52, "-m", 400, "-d", 100, "-s", 151, "-i", EXAMPLE, "-o", tmp_path / "lh_dieoff.txt"])
    run("midaspom_loss", ["-a", 10, "-e", 0.71, "-c", 0.52, "-m", 400, "-d", 100, "-s", 7, "-v", 3, "-i", EXAMPLE, "-o", tmp_path / "lh_loss.txt"])
    die = np.array([float(v) for v in (tmp_path / "lh_dieoff.txt").read_text().split()])
    loss = read_table(tmp_path / "lh_loss.txt")
    np.testing.assert_allclose(die, golden["dieoff_s151"], rtol=1e-9, atol=5.1e-21)   # reference prints %.20lf
    np.testing.assert_allclose(loss, golden["loss_s7v3"], rtol=1e-9, atol=5.1e-21)
    assert (tmp_path / "lh_dieoff.txt").read_text().count("\n") == 0 and (tmp_path / "lh_loss.txt").read_text().count("\n") == 7
    out = run("midaspom_hypothesis", [tmp_path / "lh_dieoff.txt", tmp_path / "lh_loss.txt", 8, 0.1, 100])
    got = {" ".join(l.split()[:-1]): float(l.split()[-1]) for l in out.splitlines()}
    Kd = 10 ** np.linspace(-1, 2, 151)
    i1 = int(np.argmin(np.abs(Kd - 1.0)))                       # K = 1 is a node of the 151-point grid (index 50)
    null = golden["dieoff_s151"][i1]
    want = {"AIC H0": 2 - 2 * np.log(null / 2 ** 8), "AIC H1": 4 - 2 * np.log(golden["dieoff_s151"].max() / 2 ** 8),
            "AIC H2": 6 - 2 * np.log(golden["loss_s7v3"].max() / 2 ** 8),
            "log10BF H0vsH1": np.log10(null / (golden["dieoff_s151"].sum() / 7)),
            "log10BF H0vsH2": np.log10(null / (golden["loss_s7v3"].sum() / 7 / 3)),
            "log10BF H1vsH2": np.log10(golden["dieoff_s151"].sum() / (golden["loss_s7v3"].sum() / 3))}
    for k, v in want.items():
        assert abs(got[k] - v) < 1e-6 * max(1.0, abs(v)), (k, got[k], v)
