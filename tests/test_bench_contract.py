"""bench.py contract checks that need no GPU: the reference arm's JSON line (one line on stdout, the
keys the driver reads) and the loud failure of the product arm when there is no CUDA device."""
import json
import os
import subprocess
import sys

import cases


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(cases.ROOT, "bench.py"), *args], capture_output=True,
                          text=True, timeout=600, cwd=cases.ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--workload", "sp4", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # ONE JSON line, banners go to stderr
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("LLG steps/s") and j["unit"] == "steps/s"
    assert j["higher_is_better"] is True and j["dtype"] == "f64" and j["vs_baseline"] is None
    assert j["value"] > 0 and abs(j["ms_per_step"] * j["value"] - 1e3) < 1e-6 * 1e3
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sp4" in cb["sample"]
    assert j["e2e"] == dict(value=j["value"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert j["config"]["workload"] == "sp4"
    # the WHOLE workload of the product arm (no tet-ratio extrapolation), the counts really done, and both
    # CPU figures of BASELINE.md section 3 (threaded = the line's value, serial on a sample)
    assert cb["full_workload"] is True and "the whole workload" in cb["sample"] and "186000 tets" in cb["sample"]
    assert j["steps"] == cb["steps"] == 2 and j["warmup"] == cb["warmup"] >= 1
    assert cb["serial"]["cores"] == 1 and 0 < cb["serial"]["value"] <= 1.5 * cb["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(cases.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "sp4", "--steps", "2"], capture_output=True, text=True, timeout=600,
                       cwd=cases.ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr) and "no CPU fallback" in (r.stdout + r.stderr)
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines())   # no bench line
