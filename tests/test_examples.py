"""The drop-in C++ surface: the example drivers of examples/ use only the public QuIDS API and are
built twice from the SAME source -- against the reference's headers (CPU; transcripts stored under
tests/golden/ by `make -C examples ref` + the commands in this file's docstring) and against this
repository's headers (GPU, through the C ABI).  Their transcripts must agree.

    make -C examples ref
    oracle/_ref/quantum_computer_test.ref.out > tests/golden/quantum_computer_test.txt
    oracle/_ref/qcgd_test.ref.out 6 1 > tests/golden/qcgd_test_6_1.txt
    oracle/_ref/qcgd_test.ref.out 7 5 > tests/golden/qcgd_test_7_5.txt
    oracle/_ref/quantum_computer_test.f32.ref.out > tests/golden/quantum_computer_test_f32.txt     (PROBA_TYPE = float)
    oracle/_ref/qcgd_test.f32.ref.out 6 1 > tests/golden/qcgd_test_f32_6_1.txt
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_examples():
    import quids_b200 as qb
    if not os.path.exists(qb.LIB_PATH):
        qb.build()
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", EXAMPLES, "ours"], check=True, env=env, stdout=subprocess.DEVNULL)


def normalise(text):
    """order inside a printed state is unspecified; -0 and 0 are the same amplitude"""
    blocks, current = [], []
    for line in text.splitlines():
        line = re.sub(r"(?<![\d.])-0(\.0+)?(?![\d.])", lambda m: "0" + (m.group(1) or ""), line)
        if line.startswith("\t"):
            current.append(line)
        else:
            blocks.append(sorted(current))
            current = []
            blocks.append([line])
    blocks.append(sorted(current))
    return [l for b in blocks for l in b]


def test_examples_build_against_the_drop_in_headers():
    build_examples()
    for name in ("quantum_computer_test.out", "qcgd_test.out"):
        assert os.path.exists(os.path.join(EXAMPLES, name))


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree absent")
def test_unmodified_reference_drivers_compile_against_the_drop_in_headers(tmp_path):
    """the reference's own example sources, copied verbatim into a scratch tree whose src/ is this
    repository's include/quids, compile without a single change (MPI example excluded: no MPI here)"""
    import quids_b200 as qb
    if not os.path.exists(qb.LIB_PATH):
        qb.build()
    (tmp_path / "examples").mkdir()
    os.symlink(os.path.join(ROOT, "include", "quids"), tmp_path / "src")
    for name in ("quantum_computer_test.cpp", "qcgd_test.cpp"):
        shutil.copy(os.path.join("/root/reference/examples", name), tmp_path / "examples" / name)
        out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(tmp_path / "examples" / name)],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert out.returncode == 0, out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("binary,args,golden", [("quantum_computer_test.out", [], "quantum_computer_test.txt"),
                                                ("qcgd_test.out", ["6", "1"], "qcgd_test_6_1.txt"),
                                                ("qcgd_test.out", ["7", "5"], "qcgd_test_7_5.txt")])
def test_example_transcripts_match_the_reference(binary, args, golden):
    build_examples()
    out = subprocess.run([os.path.join(EXAMPLES, binary)] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    want = open(os.path.join(GOLDEN, golden)).read()
    assert normalise(out.stdout) == normalise(want)


def parse_transcript(text, floor):
    """blocks of a driver transcript: (title without counters, {object text: amplitude} for |amplitude| >= floor, numbers of the other lines)"""
    blocks = []
    for line in text.splitlines():
        if not line.startswith("\t"):
            if line.strip():
                blocks.append([re.sub(r"\(.*?\)", "", line).strip(), {}, [float(x) for x in re.findall(r"P=([0-9.eE+-]+)", line)]])
            continue
        m = re.match(r"\t(-?[0-9.eE+-]+) ([+-]) ([0-9.eE+-]+)i  (.*)$", line)
        if m:
            z = complex(float(m.group(1)), float(m.group(3)) * (1 if m.group(2) == "+" else -1))
            if abs(z) >= floor:
                blocks[-1][1][m.group(4)] = z
        else:
            blocks[-1][2] += [float(x) for x in re.findall(r"= ([0-9.eE+-]+)", line)]
    return blocks


@pytest.mark.gpu
@pytest.mark.parametrize("binary,args,golden", [("quantum_computer_test.f32.out", [], "quantum_computer_test_f32.txt"),
                                                ("qcgd_test.f32.out", ["6", "1"], "qcgd_test_f32_6_1.txt")])
def test_float_build_of_the_drivers_matches_the_reference_float_build(binary, args, golden):
    """PROBA_TYPE = float: amplitudes agree with the reference's float build within 1e-5 (north_star's float tolerance).
    The reference's float rounding leaves residues of ~1e-8 where amplitudes cancel (objects printed as 0.00000 that the
    double arithmetic on the device removes exactly): objects below 5e-5 are outside the comparison."""
    build_examples()
    out = subprocess.run([os.path.join(EXAMPLES, binary)] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got, want = parse_transcript(out.stdout, 5e-5), parse_transcript(open(os.path.join(GOLDEN, golden)).read(), 5e-5)
    assert [b[0] for b in got] == [b[0] for b in want]
    for (title, g, gn), (_, w, wn) in zip(got, want):
        assert set(g) == set(w), title
        for k in g:
            assert abs(g[k] - w[k]) <= 2e-5, (title, k, g[k], w[k])  # 1e-5 + the 5 printed decimals
        assert len(gn) == len(wn) and all(abs(a - b) <= 1e-5 * max(1, abs(b)) for a, b in zip(gn, wn)), (title, gn, wn)
