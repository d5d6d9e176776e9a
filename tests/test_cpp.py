"""The C++ host-side mirror of the reference's classes (qec_ldpc_b200/cpp) and the init.txt-compatible CLI."""
import os
import subprocess

import numpy as np
import pytest

from util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qec_ldpc_b200")


def _env():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


@pytest.fixture(scope="module")
def host_check(qldpc, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "host_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "host_check.cpp"), "-o", exe, "-L", os.path.join(PKG, "lib"),
                    "-lqldpc_b200", "-Wl,-rpath," + os.path.join(PKG, "lib")], check=True, env=_env())
    return exe


@pytest.fixture(scope="module")
def cli(qldpc):
    from qec_ldpc_b200 import build as b
    return b.build_cli()


def test_cpp_class_surface(host_check, oracle, tmp_path):
    out = subprocess.run([host_check, str(tmp_path / "code.txt")], check=True, capture_output=True, text=True).stdout
    lines = out.splitlines()
    assert "name [J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]" in lines
    assert "dims 610 244 305 244x610 1220x1220" in lines
    assert "hHC0 1 9 20 58 34 42 12 47 57 25" in lines
    assert "syndrome_weight 4" in lines
    assert "row_is_logical 0 single_is_logical 1" in lines
    assert any(l.startswith("roundtrip 1 [J=4,K=5,L=10,P=61,s=9,t=49]") for l in lines)
    assert any(l.startswith("missing: Unable to find code file") for l in lines)
    # results-file text equals the reference's checked-in record, byte for byte (CodeStatistics.h:22-37)
    text = out.split("BEGIN_STATS\n")[1].split("END_STATS")[0]
    assert text == golden("kat_results.json")["K4a"]["text"]
    # the weight-W generator replays the reference's mt19937 stream (first frame of K4)
    x, z = oracle.weightw_stream(655687811, 30, 610, 1)
    want = " ".join("%d%s" % (v, "Y" if x[0, v] and z[0, v] else "X" if x[0, v] else "Z") for v in range(610)
                    if x[0, v] or z[0, v])
    assert "first_error " + want in lines


def test_cli_fails_loudly_without_gpu(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    (tmp_path / "init.txt").write_text("qc:3,3,6,7,2,3\n1\n1\n10\n20\n0.02")
    r = subprocess.run([cli, "init.txt", "--seed", "1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr
    assert "no usable CUDA device" in (tmp_path / "output_log.txt").read_text()


def test_cli_argument_errors(cli, tmp_path):
    r = subprocess.run([cli], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "Must provide initialization file." in (tmp_path / "output_log.txt").read_text()
    r = subprocess.run([cli, "nope.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and 'Unable to open init file "nope.txt"' in (tmp_path / "output_log.txt").read_text()


@pytest.mark.gpu
def test_cli_reproduces_reference_results_file(cli, qldpc, tmp_path):
    """Drop-in check: driven by an init file in the reference's format, the CLI appends a record identical to the
    reference's checked-in results file (K4: W=30 and W=45, 1000 frames, MAX 100, p 0.02), Duration aside."""
    from util import CODES, golden_matrix
    code = qldpc.Code.dense(*CODES["C2"], golden_matrix("C2", "pcmX"), golden_matrix("C2", "pcmZ"),
                            golden_matrix("C2", "iMinusP"))
    code.write_file(str(tmp_path / "code610.txt"))
    kat = golden("kat_results.json")
    for w, name in ((30, "K4a"), (45, "K4b")):
        (tmp_path / "init.txt").write_text("code610.txt \n%d\n%d\n1000\n100\n0.02" % (w, w))
        subprocess.run([cli, "init.txt", "--seed", str(kat[name]["seed"])], cwd=tmp_path, check=True)
        path = tmp_path / "results" / ("[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]_W_%d_MAX_100_p_0.02.txt" % w)
        got = [l for l in path.read_text().splitlines() if not l.startswith("Duration")]
        want = [l for l in kat[name]["text"].splitlines() if not l.startswith("Duration")]
        assert got[:len(want)] == want and got[len(want):] == ["", ""]
    assert "Run complete." in (tmp_path / "output_log.txt").read_text()


@pytest.mark.gpu
def test_cli_depolarizing_matches_library(cli, qldpc, tmp_path):
    (tmp_path / "init.txt").write_text("qc:4,5,10,61,9,49\n0\n0\n20000\n50\n0.05")
    subprocess.run([cli, "init.txt", "--depolarizing", "--seed", "77"], cwd=tmp_path, check=True)
    text = (tmp_path / "results" / "[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]_depolarizing_MAX_50_p_0.05.txt").read_text()
    code = qldpc.Code.qc(4, 5, 10, 61, 9, 49)
    k = qldpc.Decoder(code, 0, 20000).get_statistics_depolarizing(77, 0, 20000, 0.05, 50)["counters"]
    assert "Corrected: %d\n" % int(k[3]) in text and "Logical Errors: %d\n" % int(k[6]) in text
    assert "Errors Tested: 20000\n" in text and "Rand Seed: 77\n" in text


@pytest.mark.gpu
def test_cli_sweep_stop_rule_and_intervals(cli, qldpc, tmp_path):
    """--sweep p0:p1:k --target-errors E: every point keeps adding batches of COUNT frames of ONE continued global frame
    stream until it has seen E frame errors (or --max-frames); the record per point is in the reference's results-file
    format, the summary line carries the Wilson interval, and the counters do not depend on the device count."""
    import torch
    count, target, cap, seed, maxit = 2000, 60, 12000, 321, 50
    name = "[J=4,K=5,L=10,P=61,s=9,t=49][[n=610,k=61]]"
    code = qldpc.Code.qc(4, 5, 10, 61, 9, 49)
    dec = qldpc.Decoder(code, 0, count)
    summaries = []
    for gpus in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        d = tmp_path / ("g%d" % gpus)
        d.mkdir()
        (d / "init.txt").write_text("qc:4,5,10,61,9,49\n0\n0\n%d\n%d\n0.5" % (count, maxit))
        subprocess.run([cli, "init.txt", "--sweep", "0.03:0.07:3", "--target-errors", str(target), "--max-frames", str(cap),
                        "--seed", str(seed), "--gpus", str(gpus)], cwd=d, check=True)
        rows = [ln.split() for ln in (d / "results" / (name + "_sweep_MAX_%d.txt" % maxit)).read_text().splitlines()
                if not ln.startswith("#")]
        assert len(rows) == 3
        summaries.append(rows)
        for i, row in enumerate(rows):
            p = np.float32(0.03 + 0.04 * i / 2)
            frames, bad = int(row[1]), int(row[2])
            # stop rule: whole batches; stops at the first batch boundary with >= target errors, or at the cap
            assert frames % count == 0 and frames <= cap
            k = np.zeros(qldpc.NUM_COUNTERS, np.uint64)
            done = 0
            while True:
                k += dec.get_statistics_depolarizing(seed, (i << 40) + done, count, float(p), maxit)["counters"]
                done += count
                if done >= cap or int(k[0]) - int(k[3]) >= target:
                    break
            assert frames == done and bad == int(k[0]) - int(k[3])
            assert int(row[6]) == int(k[6]) and int(row[7]) == int(k[4]) and int(row[8]) == int(k[5])
            # Wilson 95% interval
            z, ph = 1.959963984540054, bad / frames
            den = 1 + z * z / frames
            ctr, half = (ph + z * z / (2 * frames)) / den, z * np.sqrt(ph * (1 - ph) / frames + z * z / (4.0 * frames * frames)) / den
            assert abs(float(row[4]) - (ctr - half)) < 1e-5 and abs(float(row[5]) - (ctr + half)) < 1e-5
            assert float(row[4]) <= ph <= float(row[5])
            # the per-point record in the reference's format
            text = (d / "results" / (name + "_depolarizing_MAX_%d_p_%s.txt" % (maxit, row[0]))).read_text()
            assert "Errors Tested: %d\n" % frames in text and "Corrected: %d\n" % int(k[3]) in text
            assert "Rand Seed: %d\n" % seed in text and text.startswith("Code: " + name + "\n")
    if len(summaries) == 2:
        assert summaries[0] == summaries[1]


def test_cli_rejects_wide_seed(cli, tmp_path):
    (tmp_path / "init.txt").write_text("qc:3,3,6,7,2,3\n1\n1\n10\n20\n0.02")
    r = subprocess.run([cli, "init.txt", "--depolarizing", "--seed", str(1 << 33)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "32 bits" in r.stderr
