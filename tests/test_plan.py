"""Host-side planning logic (bch_fft_b200/csrc/gm_plan.h), compiled with g++ and exercised on
the CPU: evaluation basis (Cantor chain), position permutation, pass schedule invariants."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plan_dump(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("plan") / "plan_dump")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "bch_fft_b200", "csrc"),
                    os.path.join(ROOT, "tools", "plan_dump.cpp"), "-o", exe], check=True)
    return exe


def run(exe, m, K, t, prefix=True):
    env = dict(os.environ)
    env.pop("BCHFFT_NO_CANTOR_PREFIX", None)
    if not prefix:
        env["BCHFFT_NO_CANTOR_PREFIX"] = "1"
    out = subprocess.run([exe, str(m), str(K), str(t)], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    head = dict(kv.split("=") for kv in out.stdout.splitlines()[0].split())
    passes = [dict(kv.split("=") for kv in l.split()[1:]) for l in out.stdout.splitlines()[1:]
              if not l.startswith("  pm ")]
    return {k: int(v) for k, v in head.items()}, passes


def run_pm(exe, m, t2, t1):
    """plane-major schedule lines: (kind, tile bits, window, [ops])"""
    out = subprocess.run([exe, str(m), str(m), str(t2), str(t1)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    res = []
    for l in out.stdout.splitlines():
        if l.startswith("  pm "):
            f = l.split()
            res.append((f[1], int(f[3].split("=")[1]), int(f[4].split("=")[1], 16), f[6:]))
    return res


def expected_forward_ops(m, K, t, chain):
    """op counts of gm_plan's forward stream for the periodic Cantor basis (twists only at depths that are
    positive multiples of the chain length; chain 1: every depth but the first)"""
    free = [d == 0 or (chain > 1 and d % chain != 0) for d in range(K)]
    tw = ty = tx = 0
    tcap = min(t, K)
    d = 0
    while d < K:
        tw += 0 if free[d] else 1
        run = 1
        while d + run < K and free[d + run]:
            run += 1
        h = 1
        while 2 * h <= run:
            h *= 2
        while h >= 2 and h + 1 > tcap:
            h //= 2
        if h < 2 or K - d < h:
            h = 0
        if h:
            tx += (K - h - d) + (h // 2) * (h.bit_length() - 1)
            d += h
        else:
            ty += K - 1 - d
            d += 1
    return tw, ty, tx


@pytest.mark.parametrize("m,chain", [(8, 8), (12, 4), (13, 1), (16, 16), (20, 4), (10, 2), (7, 1), (24, 8), (18, 2)])
def test_cantor_chain_length(plan_dump, m, chain):
    """Cantor chain of length 2^a for m = 2^a * odd; the periodic basis repeats it, so only the depths a, 2a, ..
    need a twist"""
    head, _ = run(plan_dump, m, m, 11, prefix=False)
    free = m if chain == m else (1 if chain == 1 else m - (m // chain - 1))
    assert head["twist_free_depths"] == free and head["perm_ok"] == 1
    assert head["tw_ops"] == m - free and head["bf_ops"] == m
    assert head["ty_ops"] == m * (m - 1) // 2 and head["tx_ops"] == 0
    # grouped Cantor conversion (gm_plan.h): runs of h depths without a twist between them
    head, _ = run(plan_dump, m, m, 11)
    tw, ty, tx = expected_forward_ops(m, m, 11, chain)
    assert (head["tw_ops"], head["ty_ops"], head["tx_ops"], head["bf_ops"]) == (tw, ty, tx, m)


@pytest.mark.parametrize("m,K,t", [(16, 16, 11), (20, 20, 11), (13, 13, 12), (13, 5, 11), (16, 10, 11),
                                   (10, 10, 4), (8, 3, 2), (20, 7, 11)])
def test_pass_schedule_invariants(plan_dump, m, K, t):
    head, passes = run(plan_dump, m, K, t)
    assert passes
    nops = 0
    for p in passes:
        dom, tt, window = int(p["dom"]), int(p["t"]), int(p["window"], 16)
        assert tt == min(t, dom) and bin(window).count("1") == tt and window < (1 << max(dom, 1))
        # the window is at most two runs of bits
        assert len(re.findall(r"1+", bin(window)[2:])) <= 2
        nops += int(p["ops"])
    assert nops == head["tw_ops"] + head["ty_ops"] + head["bf_ops"] + head["tx_ops"]
    assert head["bf_ops"] == K


def test_gf2_16_is_nine_passes(plan_dump):
    head, passes = run(plan_dump, 16, 16, 11, prefix=False)
    assert head["passes"] == 9 and head["tw_ops"] == 0
    # generic tile schedule with the grouped Cantor conversion: five passes; GF(2^20): 9 instead of 17
    head, passes = run(plan_dump, 16, 16, 11)
    assert head["passes"] == 5 and head["tw_ops"] == 0
    assert run(plan_dump, 20, 20, 11, prefix=False)[0]["passes"] == 17      # plain Taylor levels, no required bit
    head20, _ = run(plan_dump, 20, 20, 11)
    assert head20["passes"] == 9 and head20["tw_ops"] == 4 and head20["tx_ops"] == 60 and head20["ty_ops"] == 0


def test_plane_major_schedule_gf2_16(plan_dump):
    """Cantor basis conversion: m/2 * log2(m) = 32 generalised Taylor levels instead of 120: one
    line pass (levels whose super-block exceeds a tile), one one-plane tile pass, the rest folded
    into the first butterfly pass; every tile window keeps position bits 0..2 (32-byte runs per
    plane row)"""
    pm = run_pm(plan_dump, 16, 11, 14)
    assert [k for k, *_ in pm] == ["line", "plane", "bf", "bf"]
    assert pm[0][3] == ["TX(7,8)", "TX(6,8)"]
    assert [k for k, *_ in run_pm(plan_dump, 16, 11, 13)] == ["line", "plane", "bf", "bf"]
    ops = [o for *_, oo in pm for o in oo]
    assert sum(o.startswith("TX") for o in ops) == 32 and sum(o.startswith("BF") for o in ops) == 16
    assert ops[-16:] == [f"BF({d},0)" for d in range(15, -1, -1)]
    for kind, t, window, oo in pm:
        assert bin(window).count("1") == t and window & 7 == 7
        assert len(re.findall(r"1+", bin(window)[2:])) <= 2
        for o in oo:
            a, b = (int(x) for x in o[3:-1].split(","))
            bits = (((2 << b) - 1) << a) if o.startswith("TX") else (1 << a)
            assert bits & ~window == 0, (o, hex(window))


def test_plane_major_needs_a_power_of_two_degree(plan_dump):
    assert run_pm(plan_dump, 20, 11, 14) == [] and run_pm(plan_dump, 13, 11, 14) == []
    assert [k for k, *_ in run_pm(plan_dump, 8, 6, 8)] == ["plane", "bf", "bf"]
    assert [k for k, *_ in run_pm(plan_dump, 8, 6, 6)] == ["line", "plane", "plane", "plane", "bf", "bf"]
