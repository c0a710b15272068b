"""The C-ABI as a C consumer sees it: include/b200rank.h compiles as strict C99 and as C++, examples/score_yes_no.c (a plain-C host:
create -> load_tensor by HF name -> score_yes_no -> destroy) builds against libb200rank.so, fails loudly without a GPU, and — on
the B200 — prints the numbers the CPU oracle computes for the same model."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, c_example_model

LIBDIR = os.path.join(ROOT, "llm-rankers_b200")
EXAMPLE = os.path.join(ROOT, "examples", "score_yes_no.c")


def build_example(tmp_path):
    exe = str(tmp_path / "score_yes_no")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), EXAMPLE,
                           "-L", LIBDIR, "-lb200rank", f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", exe])
    return exe


def parse(text):
    rows = re.findall(r"doc (\d+): yes (-?[\d.]+) no (-?[\d.]+) P\(yes\) ([\d.]+)", text)
    assert [int(r[0]) for r in rows] == list(range(len(rows))) and rows
    return np.array([[float(r[1]), float(r[2])] for r in rows]), np.array([float(r[3]) for r in rows])


def oracle_answers():
    from oracle.t5_oracle import T5Oracle
    cfg, w, ids, lengths = c_example_model()
    mask = (np.arange(ids.shape[1])[None] < lengths[:, None]).astype(np.int64)
    return T5Oracle(cfg, w).score_yes_no(ids.astype(np.int64), mask, 12, 13)


def check_against_oracle(text):
    lg, sc = parse(text)
    ref_lg, ref_sc = oracle_answers()
    from b200rank.tolerance import logit_tolerance
    assert np.all(np.abs(lg - ref_lg) <= logit_tolerance(ref_lg, 0)), (lg, ref_lg)
    assert np.abs(sc - ref_sc).max() < 0.01


@pytest.mark.parametrize("compiler,std", [("gcc", "-std=c99"), ("gcc", "-std=c11"), ("g++", "-std=c++11")])
def test_header_is_self_contained(compiler, std, tmp_path):
    src = tmp_path / ("t.c" if compiler == "gcc" else "t.cpp")
    src.write_text('#include "b200rank.h"\nint main(void) { b200rank_config c; (void)c; return B200RANK_OK; }\n')
    subprocess.check_call([compiler, std, "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


@pytest.mark.skipif(__import__("conftest").HAS_GPU, reason="checks the no-GPU failure mode")
def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    exe = build_example(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)
    assert "b200rank" in p.stdout and "no CPU fallback" in p.stderr


def test_committed_b200_output_of_the_c_example_matches_the_oracle():
    """tests/golden/c_example_output_b200.txt is what the example printed on a B200 (gpurun, round 1); the model is regenerated here from
    the example's own xorshift stream (helpers.c_example_model), so this pins both the fixture and that restatement."""
    with open(os.path.join(GOLDEN, "c_example_output_b200.txt")) as f:
        check_against_oracle(f.read())
