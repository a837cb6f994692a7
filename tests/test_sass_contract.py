"""Post-build SASS contract (no GPU): the march loops consume tcgen05.ld results without a tcgen05.wait::ld
statement (tmem.cuh, KW_TMEM_HOT_WAIT = 0; 4 % faster than with the statement, profiles/r2_f_*).  PTX asks for the wait;
what makes the code correct is that ptxas gives every LDTM a write scoreboard and puts the wait on the first
consumer.  tools/sass_contract.py re-derives that from the built library's machine code, so a toolchain that stops
doing it fails this test instead of silently mispricing."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
LIB = os.path.join(ROOT, "kwinto-cuda_b200", "lib", "libkwfd1d.so")


@pytest.fixture(scope="module")
def contract():
    import shutil
    import subprocess

    import sass_contract

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(LIB):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "kwinto-cuda_b200", "csrc")], check=True)
    sass_contract.check_function.warnings = []
    report, problems = sass_contract.run(LIB, verbose=False)
    return sass_contract, report, problems


def test_every_ldtm_is_scoreboarded_and_waited_for(contract):
    _, report, problems = contract
    assert problems == [], "\n".join(problems[:20])
    by_name = {name: (n_ldtm, n_sttm, march) for name, _, n_ldtm, n_sttm, march in report}
    # the kernels that keep a~, g~, D, p in tensor memory really do (SASS: LDTM / STTM), and their march loops were found
    for key, min_loops in (("fd1d_iw_kernelILi4ELi2ELb0ELb0ELi1Ed", 5), ("fd1d_iw_kernelILi4ELi2ELb1ELb0ELi1Ed", 10), ("fd1d_iw_kernelILi4ELi2ELb0ELb0ELi2Ed", 4), ("fd1d_iw_kernelILi4ELi2ELb0ELb0ELi4Ed", 3), ("fd1d_wide_kernelILi4", 5),
                           ("fd1d_wide_kernelILi2", 5)):
        hits = [v for k, v in by_name.items() if key in k]
        assert hits, key
        for n_ldtm, n_sttm, march in hits:
            assert n_ldtm >= 16 and n_sttm >= 4, (key, n_ldtm, n_sttm)
            assert len(march) >= min_loops, (key, len(march))
    # the headline kernel: every march loop has 12 LDTM.x32 per step (a~, g~, D, p of a chunk pair + a~, g~ again, two
    # pairs) and at most 16 moves
    (n_ldtm, n_sttm, march), = [v for k, v in by_name.items() if "fd1d_iw_kernelILi4ELi2ELb0ELb0ELi1Ed" in k]
    steps = [m for a, b, m in march if m["LDTM"] == 12 and m["DSETP"] == 32]  # the five level-specialised march steps
    assert len(steps) == 5
    for mix in steps:
        assert mix["LDTM"] == 12 and mix["DFMA"] >= 170 and mix["IMAD"] + mix["MOV"] <= 16, dict(mix)
        assert mix["LDL"] == 0 and mix["STL"] == 0 and mix["LDS"] == 0


def test_the_checker_sees_a_missing_wait(contract):
    """Negative control: strip the scoreboard waits (or the scoreboards) from the parsed headline kernel and the
    checker must object."""
    import subprocess

    sc, report, _ = contract
    (mangled,) = [name for name, *_ in report if "fd1d_iw_kernelILi4ELi2ELb0ELb0ELi1Ed" in name]  # the headline kernel
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, LIB], capture_output=True, text=True, check=True).stdout
    (name, body), = sc.split_functions(sass).items()
    code = sc.parse_function(body)
    assert sc.check_function(name, code)[2] == []
    for ins in code:
        ins.wait = 0
    assert len(sc.check_function(name, code)[2]) > 100
    code = sc.parse_function(body)
    for ins in code:
        if ins.base == "LDTM":
            ins.wbar = 7
    assert len(sc.check_function(name, code)[2]) > 100


def test_committed_hot_loop_listing():
    """profiles/ keeps the SASS of the headline march loop with its LDTM / DFMA lines and control words."""
    path = os.path.join(ROOT, "profiles", "r2_sass_hotloop_v237_level2.txt")
    text = open(path).read()
    assert text.count("LDTM.x32") == 12 and text.count("DFMA") >= 170 and "wait 000001" in text
